/*
 * bvcf_oracle.c -- CPU ORACLE (test infrastructure, never shipped; see bvcf_oracle.h).
 *
 * Plain-C restatement of /root/reference/main.go for the per-line VCF transform.
 * Every function cites the main.go lines it follows.  Parity pinned against the
 * reference golden + main_test.go vectors (tests/test_oracle_*.py).
 */
#define _GNU_SOURCE
#include "bvcf_oracle.h"

#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ---------- small utilities ---------- */

typedef struct {
  char *p;
  size_t n, cap;
} obuf;

static void ob_reserve(obuf *b, size_t extra) {
  if (b->n + extra <= b->cap) return;
  size_t c = b->cap ? b->cap : 4096;
  while (c < b->n + extra) c *= 2;
  b->p = (char *)realloc(b->p, c);
  b->cap = c;
}
static void ob_put(obuf *b, const char *s, size_t n) {
  ob_reserve(b, n);
  memcpy(b->p + b->n, s, n);
  b->n += n;
}
static void ob_putc(obuf *b, char c) {
  ob_reserve(b, 1);
  b->p[b->n++] = c;
}
static void ob_puts(obuf *b, const char *s) { ob_put(b, s, strlen(s)); }

typedef struct {
  const char *p;
  size_t n;
} sv; /* string view */

static int sv_eq(sv a, const char *s) {
  size_t n = strlen(s);
  return a.n == n && memcmp(a.p, s, n) == 0;
}

/* strconv.Itoa */
static size_t itoa64(long long v, char *out) {
  char tmp[24];
  int k = 0;
  unsigned long long u = v < 0 ? 0ULL - (unsigned long long)v : (unsigned long long)v;
  do {
    tmp[k++] = (char)('0' + u % 10);
    u /= 10;
  } while (u);
  size_t n = 0;
  if (v < 0) out[n++] = '-';
  while (k) out[n++] = tmp[--k];
  out[n] = 0;
  return n;
}

/* strconv.Atoi (64-bit int): optional sign, >=1 decimal digits, no overflow. Returns 0 on success. */
static int go_atoi(sv s, long long *out) {
  size_t i = 0;
  int neg = 0;
  if (s.n == 0) return -1;
  if (s.p[0] == '-' || s.p[0] == '+') {
    neg = s.p[0] == '-';
    i = 1;
    if (s.n == 1) return -1;
  }
  unsigned long long v = 0;
  for (; i < s.n; i++) {
    unsigned d = (unsigned char)s.p[i] - '0';
    if (d > 9) return -1;
    if (v > (0xFFFFFFFFFFFFFFFFULL - d) / 10) return -1;
    v = v * 10 + d;
  }
  if (!neg && v > 0x7FFFFFFFFFFFFFFFULL) return -1;
  if (neg && v > 0x8000000000000000ULL) return -1;
  *out = neg ? (long long)(0ULL - v) : (long long)v;
  return 0;
}

/* ---------- bystro-utils/parse symbols (values pinned by main_test.go:79-80 + golden) ---------- */

static const char *const PARSE_HEADER[15] = {
    "chrom", "pos", "type", "ref", "alt", "trTv", "heterozygotes", "heterozygosity",
    "homozygotes", "homozygosity", "missingGenos", "missingness", "ac", "an", "sampleMaf"};
static const char SNP[] = "SNP", INS[] = "INS", DEL[] = "DEL", MNP[] = "MNP", MULTI[] = "MULTIALLELIC";

/* parse.GetTrTv (main.go:605): non single-base alt => "0"; A<->G, C<->T => "1"; else "2" */
const char *oracle_tr_tv(char ref, const char *alt, size_t alt_len) {
  if (alt_len != 1) return "0";
  char a = alt[0];
  if ((ref == 'A' && a == 'G') || (ref == 'G' && a == 'A') || (ref == 'C' && a == 'T') ||
      (ref == 'T' && a == 'C'))
    return "1";
  return "2";
}

/* strconv.FormatFloat(q,'G',3,64) == C "%.3G" on (0,1]; glibc rounds the exact
 * binary value half-to-even, strips trailing zeros and prints a 2-digit exponent like Go. */
void oracle_format_float(double q, char out[32]) { snprintf(out, 32, "%.3G", q); }

/* main.go:219-239 */
char *oracle_header(const oracle_config *cfg) {
  obuf b = {0};
  for (int i = 0; i < 15; i++) {
    if (i) ob_putc(&b, '\t');
    ob_puts(&b, PARSE_HEADER[i]);
  }
  if (cfg->keep_pos) ob_puts(&b, "\tvcfPos");
  if (cfg->keep_id) ob_puts(&b, "\tid");
  if (cfg->keep_info) ob_puts(&b, "\talleleIdx\tinfo");
  ob_putc(&b, 0);
  return b.p;
}

/* main.go:456-474 */
int oracle_alt_is_valid(const char *alt, size_t n) {
  if (n == 0) return 0; /* reference panics on alt[0]; out of contract => invalid */
  for (size_t i = 0; i < n; i++)
    if (alt[i] != 'A' && alt[i] != 'C' && alt[i] != 'T' && alt[i] != 'G') return 0;
  return 1;
}

/* ---------- getAlleles (main.go:723-1038) ---------- */

typedef struct {
  char *pos; /* malloc'd: a verbatim POS field is copied whole, whatever its length (main.go:742,803) */
  char ref;
  char *alt; /* malloc'd */
  size_t alt_len;
  int idx;
} out_allele;

typedef struct {
  out_allele *a;
  int n, cap;
  const char *type;
} allele_list;

typedef struct {
  oracle_diag *d;
  size_t n, cap;
  uint64_t line_no;
} diag_sink;

static void diag(diag_sink *s, int alt_no, int code) {
  if (!s) return;
  if (s->n == s->cap) {
    s->cap = s->cap ? s->cap * 2 : 64;
    s->d = (oracle_diag *)realloc(s->d, s->cap * sizeof(oracle_diag));
  }
  s->d[s->n].line_no = s->line_no;
  s->d[s->n].alt_no = alt_no;
  s->d[s->n].code = code;
  s->n++;
}

static void al_push(allele_list *l, const char *pos, size_t pos_n, char ref, const char *pfx,
                    const char *alt, size_t alt_n, int idx) {
  if (l->n == l->cap) {
    l->cap = l->cap ? l->cap * 2 : 8;
    l->a = (out_allele *)realloc(l->a, l->cap * sizeof(out_allele));
  }
  out_allele *o = &l->a[l->n++];
  o->pos = (char *)malloc(pos_n + 1);
  memcpy(o->pos, pos, pos_n);
  o->pos[pos_n] = 0;
  o->ref = ref;
  size_t pn = strlen(pfx);
  o->alt = (char *)malloc(pn + alt_n + 1);
  memcpy(o->alt, pfx, pn);
  memcpy(o->alt + pn, alt, alt_n);
  o->alt[pn + alt_n] = 0;
  o->alt_len = pn + alt_n;
  o->idx = idx;
}
static void al_push_int(allele_list *l, long long pos, char ref, long long alt, int idx) {
  char pb[24], ab[24];
  size_t pn = itoa64(pos, pb);
  size_t an = itoa64(alt, ab);
  al_push(l, pb, pn, ref, "", ab, an, idx);
}
static void al_free(allele_list *l) {
  for (int i = 0; i < l->n; i++) { free(l->a[i].alt); free(l->a[i].pos); }
  free(l->a);
  l->a = NULL;
  l->n = l->cap = 0;
}

static void get_alleles(sv pos, sv ref, sv alt, allele_list *out, diag_sink *ds) {
  out->n = 0;
  out->type = "";
  /* reference panics on empty REF/ALT (main.go:747); out of contract => no alleles */
  if (ref.n == 0 || alt.n == 0) return;

  if (alt.n == ref.n && memcmp(alt.p, ref.p, alt.n) == 0) { /* :729 */
    diag(ds, 0, ORACLE_DIAG_SAME);
    return;
  }
  if (alt.n == 1) { /* :735 */
    char a = alt.p[0];
    if (a != 'A' && a != 'C' && a != 'G' && a != 'T') {
      diag(ds, 1, ORACLE_DIAG_BAD_ALT);
      return;
    }
    if (ref.n == 1) { /* :742 */
      al_push(out, pos.p, pos.n, ref.p[0], "", alt.p, 1, 0);
      out->type = SNP;
      return;
    }
    if (a != ref.p[0]) { /* :747 */
      diag(ds, 1, ORACLE_DIAG_DEL1);
      return;
    }
    long long ip;
    if (go_atoi(pos, &ip)) { /* :752 */
      diag(ds, 1, ORACLE_DIAG_POS);
      return;
    }
    al_push_int(out, ip + 1, ref.p[1], 1 - (long long)ref.n, 0); /* :764 */
    out->type = DEL;
    return;
  }

  long long ip = 0;
  int multi = 0;
  int alt_idx = 0;
  size_t s = 0;
  for (;; alt_idx++) { /* strings.Split(alt, ",") :774 */
    size_t e = s;
    while (e < alt.n && alt.p[e] != ',') e++;
    sv t = {alt.p + s, e - s};
    int last = (e >= alt.n);
    s = e + 1;

    if (alt_idx > 0) multi = 1; /* :777 */

    do {
      if (!oracle_alt_is_valid(t.p, t.n)) { /* :781 */
        diag(ds, alt_idx + 1, ORACLE_DIAG_BAD_ALT);
        break;
      }
      if (ref.n == 1) {   /* :786 */
        if (t.n == 1) { /* :787 */
          al_push(out, pos.p, pos.n, ref.p[0], "", t.p, 1, alt_idx);
          break;
        }
        if (t.p[0] != ref.p[0]) { /* :797 */
          diag(ds, alt_idx + 1, ORACLE_DIAG_INS1);
          break;
        }
        al_push(out, pos.p, pos.n, ref.p[0], "+", t.p + 1, t.n - 1, alt_idx); /* :803-812 */
        break;
      }
      if (ip == 0) { /* :822 */
        if (go_atoi(pos, &ip)) {
          diag(ds, 0, ORACLE_DIAG_POS_LIST);
          last = 2; /* break out of the allele loop, keep what we have :828 */
          break;
        }
      }
      if (t.n == 1) { /* :832 */
        if (t.p[0] != ref.p[0]) {
          diag(ds, alt_idx + 1, ORACLE_DIAG_DEL1_LIST);
          break;
        }
        al_push_int(out, ip + 1, ref.p[1], 1 - (long long)ref.n, alt_idx);
        break;
      }
      if (ref.n == t.n) { /* :855 MNP / padded SNP */
        for (size_t i = 0; i < ref.n; i++) {
          if (ref.p[i] != t.p[i]) {
            char pb[24];
            size_t pn = itoa64(ip + (long long)i, pb);
            al_push(out, pb, pn, ref.p[i], "", t.p + i, 1, alt_idx);
          }
        }
        break;
      }
      if (t.n > ref.n) { /* :899 insertion */
        long long r = 0;
        while ((long long)t.n + r > 0 && (long long)ref.n + r > 1 &&
               t.p[(long long)t.n + r - 1] == ref.p[(long long)ref.n + r - 1])
          r--;
        long long off = (long long)ref.n + r; /* :932 */
        if (memcmp(ref.p, t.p, (size_t)off) != 0) {
          diag(ds, alt_idx + 1, ORACLE_DIAG_MIXED);
          break;
        }
        char pb[24];
        size_t pn = itoa64(ip + off - 1, pb);
        al_push(out, pb, pn, ref.p[off - 1], "+", t.p + off, (size_t)((long long)t.n + r - off), alt_idx);
        break;
      }
      { /* :971 deletion */
        long long r = 0;
        while ((long long)t.n + r > 1 && (long long)ref.n + r > 0 &&
               t.p[(long long)t.n + r - 1] == ref.p[(long long)ref.n + r - 1])
          r--;
        long long off = (long long)t.n + r; /* :984 */
        if (memcmp(ref.p, t.p, (size_t)off) != 0) {
          diag(ds, alt_idx + 1, ORACLE_DIAG_MIXED);
          break;
        }
        al_push_int(out, ip + off, ref.p[off], -((long long)ref.n + r - off), alt_idx);
      }
    } while (0);

    if (last) break;
  }

  if (out->n == 0) return; /* :1004 */
  if (multi) {
    out->type = MULTI;
    return;
  }
  if (out->a[0].alt_len > 1) { /* :1018 */
    out->type = out->a[0].alt[0] == '-' ? DEL : INS;
    return;
  }
  out->type = out->n > 1 ? MNP : SNP; /* :1032-1037 */
}

int oracle_get_alleles(const char *chrom, const char *pos, const char *ref, const char *alt,
                       char *type_out, int cap, char (*pos_out)[24], char *ref_out, char **alt_out,
                       int *idx_out) {
  (void)chrom;
  allele_list l = {0};
  sv p = {pos, strlen(pos)}, r = {ref, strlen(ref)}, a = {alt, strlen(alt)};
  get_alleles(p, r, a, &l, NULL);
  strcpy(type_out, l.type);
  int n = l.n;
  for (int i = 0; i < n && i < cap; i++) {
    snprintf(pos_out[i], 24, "%s", l.a[i].pos); /* unit-level API: 23 characters are plenty for the vectors */
    ref_out[i] = l.a[i].ref;
    alt_out[i] = strdup(l.a[i].alt);
    idx_out[i] = l.a[i].idx;
  }
  al_free(&l);
  return n;
}

/* ---------- makeHetHomozygotes (main.go:1042-1194), general form == fast path ---------- */

/* classify one sample field for allele string a. Returns 0 none, 1 het, 2 hom, 3 missing.
 * *gt / *alt get the counts that are added to an / ac (0 when missing). */
static int classify_sample(sv f, sv a, int *gt_out, int *alt_out) {
  /* SplitN(field, ":", 2)[0]  :1127 */
  size_t gn = 0;
  while (gn < f.n && f.p[gn] != ':') gn++;
  char sep = 0; /* :1130-1137 */
  if (memchr(f.p, '|', gn)) sep = '|';
  else if (memchr(f.p, '/', gn)) sep = '/';
  int gt = 0, alt = 0;
  size_t s = 0;
  for (;;) { /* :1149 */
    size_t e = s;
    if (sep) {
      while (e < gn && f.p[e] != sep) e++;
    } else {
      e = gn;
    }
    size_t tn = e - s;
    if (tn == 1 && f.p[s] == '.') { /* :1150 */
      *gt_out = 0;
      *alt_out = 0;
      return 3;
    }
    if (tn == a.n && memcmp(f.p + s, a.p, tn) == 0) alt++; /* :1162 */
    gt++;
    if (e >= gn) break;
    s = e + 1;
  }
  *gt_out = gt;
  *alt_out = alt;
  if (alt == 0) return 0;
  return alt == gt ? 2 : 1; /* :1185 */
}

void oracle_het_hom(const char *const *fields, int n, const char *allele_num, uint8_t *flags,
                    int8_t *dosage, int *ac, int *an) {
  sv a = {allele_num, strlen(allele_num)};
  int tac = 0, tan = 0;
  for (int i = 0; i < n; i++) {
    sv f = {fields[i], strlen(fields[i])};
    int gt, alt;
    int c = classify_sample(f, a, &gt, &alt);
    flags[i] = (uint8_t)c;
    tac += alt;
    tan += gt;
    if (dosage) dosage[i] = c == 3 ? -1 : (int8_t)(alt <= 127 ? alt : 127); /* :1172-1178 */
  }
  *ac = tac;
  *an = tan;
}

/* ---------- processLines (main.go:476-721) ---------- */

typedef struct {
  const oracle_config *cfg;
  sv *header;    /* header fields (sample names normalised) */
  int n_header;
  int eol_width; /* numChars */
} stream_ctx;

typedef struct {
  obuf tsv;
  obuf loci;
  obuf dosage;
  uint64_t n_rows, n_dosage_rows;
  diag_sink ds;
  /* scratch */
  sv *rec;
  uint8_t *flags;
  int8_t *dos;
} worker;

static int in_set(const char *const *set, int n, sv v) {
  for (int i = 0; i < n; i++)
    if (sv_eq(v, set[i])) return 1;
  return 0;
}

static void put_names(obuf *o, const stream_ctx *sc, const uint8_t *flags, int cls, const char *delim,
                      size_t dl) {
  int first = 1;
  int ns = sc->n_header - 9;
  for (int i = 0; i < ns; i++) {
    if (flags[i] != cls) continue;
    if (!first) ob_put(o, delim, dl);
    first = 0;
    ob_put(o, sc->header[9 + i].p, sc->header[9 + i].n);
  }
}

static void process_line(const stream_ctx *sc, worker *w, const char *row, size_t row_len) {
  const oracle_config *cfg = sc->cfg;
  /* strings.Split(string(row[:len(row)-numChars]), "\t")  :535 */
  size_t n = row_len >= (size_t)sc->eol_width ? row_len - (size_t)sc->eol_width : 0;
  int nf = 0;
  size_t s = 0;
  int H = sc->n_header;
  for (;;) {
    const char *t = (const char *)memchr(row + s, '\t', n - s);
    size_t e = t ? (size_t)(t - row) : n;
    if (nf < H) {
      w->rec[nf].p = row + s;
      w->rec[nf].n = e - s;
    }
    nf++;
    if (!t) break;
    s = e + 1;
  }
  /* linePasses :447-454 */
  if (nf != H) return;
  if (H < 8) return; /* reference would panic indexing record[filterIdx]; out of contract */
  if (cfg->n_allow >= 0 && !in_set(cfg->allow, cfg->n_allow, w->rec[6])) return;
  if (cfg->n_exclude > 0 && in_set(cfg->exclude, cfg->n_exclude, w->rec[6])) return;

  allele_list al = {0};
  get_alleles(w->rec[1], w->rec[3], w->rec[4], &al, &w->ds); /* :541 */
  if (al.n == 0) {
    al_free(&al);
    return;
  }
  int multiallelic = al.type == MULTI; /* :547 */
  int n_samples = H > 9 ? H - 9 : 0;   /* :505-509 */
  size_t dl = strlen(cfg->field_delim), el = strlen(cfg->empty_field);
  sv chrom = w->rec[0];
  int add_chr = chrom.n < 4 || chrom.p[0] != 'c'; /* :570 */

  for (int i = 0; i < al.n; i++) { /* :549 */
    int ac = 0, an = 0, n_het = 0, n_hom = 0, n_miss = 0;
    double eff = 0;
    char num[24];
    size_t numn = itoa64(al.a[i].idx + 1, num); /* :552 */
    if (n_samples > 0) {
      sv a = {num, numn};
      for (int k = 0; k < n_samples; k++) { /* makeHetHomozygotes :556 */
        int gt, alt;
        int c = classify_sample(w->rec[9 + k], a, &gt, &alt);
        w->flags[k] = (uint8_t)c;
        ac += alt;
        an += gt;
        n_het += c == 1;
        n_hom += c == 2;
        n_miss += c == 3;
        if (cfg->want_dosage) w->dos[k] = c == 3 ? -1 : (int8_t)(alt <= 127 ? alt : 127);
      }
      if (ac == 0) continue; /* :558 */
      /* needsLabels==false leaves `missing` empty (:1114), so effectiveSamples == numSamples then */
      eff = (double)n_samples - (cfg->want_tsv ? (double)n_miss : 0.0); /* :563 */
    }

    if (cfg->want_dosage && n_samples > 0) { /* :576-584; no arrow writer when there are no samples :308-318 */
      if (add_chr) ob_puts(&w->loci, "chr");
      ob_put(&w->loci, chrom.p, chrom.n);
      ob_putc(&w->loci, ':');
      ob_puts(&w->loci, al.a[i].pos);
      ob_putc(&w->loci, ':');
      ob_putc(&w->loci, al.a[i].ref);
      ob_putc(&w->loci, ':');
      ob_put(&w->loci, al.a[i].alt, al.a[i].alt_len);
      ob_putc(&w->loci, '\n');
      ob_put(&w->dosage, (const char *)w->dos, (size_t)n_samples);
      w->n_dosage_rows++;
    }

    if (!cfg->want_tsv) continue; /* needsLabels :586 */
    obuf *o = &w->tsv;
    if (add_chr) ob_puts(o, "chr");
    ob_put(o, chrom.p, chrom.n);
    ob_putc(o, '\t');
    ob_puts(o, al.a[i].pos);
    ob_putc(o, '\t');
    ob_puts(o, al.type);
    ob_putc(o, '\t');
    ob_putc(o, al.a[i].ref);
    ob_putc(o, '\t');
    ob_put(o, al.a[i].alt, al.a[i].alt_len);
    ob_putc(o, '\t');
    ob_puts(o, multiallelic ? "0" : oracle_tr_tv(al.a[i].ref, al.a[i].alt, al.a[i].alt_len)); /* :602-606 */
    ob_putc(o, '\t');
    char fb[32];
    /* hets :612-628 */
    if (n_het == 0) {
      ob_put(o, cfg->empty_field, el);
      ob_puts(o, "\t0");
    } else {
      put_names(o, sc, w->flags, 1, cfg->field_delim, dl);
      ob_putc(o, '\t');
      oracle_format_float((double)n_het / eff, fb);
      ob_puts(o, fb);
    }
    ob_putc(o, '\t');
    /* homs :634-642 */
    if (n_hom == 0) {
      ob_put(o, cfg->empty_field, el);
      ob_puts(o, "\t0");
    } else {
      put_names(o, sc, w->flags, 2, cfg->field_delim, dl);
      ob_putc(o, '\t');
      oracle_format_float((double)n_hom / eff, fb);
      ob_puts(o, fb);
    }
    ob_putc(o, '\t');
    /* missing :648-656 */
    if (n_miss == 0) {
      ob_put(o, cfg->empty_field, el);
      ob_puts(o, "\t0");
    } else {
      put_names(o, sc, w->flags, 3, cfg->field_delim, dl);
      ob_putc(o, '\t');
      oracle_format_float((double)n_miss / (double)n_samples, fb);
      ob_puts(o, fb);
    }
    ob_putc(o, '\t');
    char ib[24];
    ob_put(o, ib, itoa64(ac, ib)); /* :661 */
    ob_putc(o, '\t');
    ob_put(o, ib, itoa64(an, ib)); /* :663 */
    ob_putc(o, '\t');
    if (ac == 0) { /* :667 */
      ob_putc(o, '0');
    } else {
      oracle_format_float((double)ac / (double)an, fb);
      ob_puts(o, fb);
    }
    if (cfg->keep_pos) { /* :674 */
      ob_putc(o, '\t');
      ob_put(o, w->rec[1].p, w->rec[1].n);
    }
    if (cfg->keep_id) { /* :679 */
      ob_putc(o, '\t');
      ob_put(o, w->rec[2].p, w->rec[2].n);
    }
    if (cfg->keep_info) { /* :684 */
      ob_putc(o, '\t');
      ob_put(o, ib, itoa64(al.a[i].idx, ib));
      ob_putc(o, '\t');
      ob_put(o, w->rec[7].p, w->rec[7].n);
    }
    ob_putc(o, '\n');
    w->n_rows++;
  }
  al_free(&al);
}

typedef struct {
  const stream_ctx *sc;
  worker w;
  const char *data;
  size_t len;
  uint64_t first_line;
  uint64_t n_lines;
  char eol;
} job;

static void *run_job(void *arg) {
  job *j = (job *)arg;
  const stream_ctx *sc = j->sc;
  int H = sc->n_header;
  j->w.rec = (sv *)calloc((size_t)(H > 0 ? H : 1), sizeof(sv));
  int ns = H > 9 ? H - 9 : 0;
  j->w.flags = (uint8_t *)calloc((size_t)ns + 1, 1);
  j->w.dos = (int8_t *)calloc((size_t)ns + 1, 1);
  size_t s = 0;
  uint64_t ln = j->first_line;
  while (s < j->len) { /* ReadBytes(eol) :354 */
    const char *nl = (const char *)memchr(j->data + s, j->eol, j->len - s);
    if (!nl) break; /* io.EOF: unterminated last line is dropped :356 */
    size_t e = (size_t)(nl - j->data) + 1;
    j->w.ds.line_no = ln++;
    process_line(sc, &j->w, j->data + s, e - s);
    s = e;
  }
  j->n_lines = ln - j->first_line;
  free(j->w.rec);
  free(j->w.flags);
  free(j->w.dos);
  return NULL;
}

static int parse_header_line(const oracle_config *cfg, const char *line, size_t n, sv **hdr, int *nh,
                             char **owned) {
  /* copy so that NormalizeHeader can mutate :296 */
  char *c = (char *)malloc(n + 1);
  memcpy(c, line, n);
  c[n] = 0;
  int cnt = 1;
  for (size_t i = 0; i < n; i++) cnt += c[i] == '\t';
  sv *h = (sv *)calloc((size_t)cnt, sizeof(sv));
  int k = 0;
  size_t s = 0;
  for (size_t i = 0; i <= n; i++) {
    if (i == n || c[i] == '\t') {
      h[k].p = c + s;
      h[k].n = i - s;
      k++;
      s = i + 1;
    }
  }
  if (cfg->normalize_dots)
    for (int f = 9; f < cnt; f++)
      for (size_t i = 0; i < h[f].n; i++)
        if (h[f].p[i] == '.') ((char *)h[f].p)[i] = '_';
  *hdr = h;
  *nh = cnt;
  *owned = c;
  return 0;
}

static int run_block(const oracle_config *cfg, sv *hdr, int nh, int eol_width, char eol, const char *block,
                     size_t len, int threads, oracle_result *res) {
  stream_ctx sc = {cfg, hdr, nh, eol_width};
  if (threads < 1) threads = 1;
  job *jobs = (job *)calloc((size_t)threads, sizeof(job));
  /* split into contiguous newline-aligned blocks */
  size_t start = 0;
  int nj = 0;
  for (int t = 0; t < threads; t++) {
    size_t end = t == threads - 1 ? len : (len / (size_t)threads) * (size_t)(t + 1);
    if (end < start) end = start;
    if (t != threads - 1) {
      const char *nl = end < len ? (const char *)memchr(block + end, eol, len - end) : NULL;
      end = nl ? (size_t)(nl - block) + 1 : len;
    }
    jobs[nj].sc = &sc;
    jobs[nj].data = block + start;
    jobs[nj].len = end - start;
    jobs[nj].eol = eol;
    nj++;
    start = end;
  }
  if (nj == 1) {
    run_job(&jobs[0]);
  } else {
    pthread_t *th = (pthread_t *)calloc((size_t)nj, sizeof(pthread_t));
    for (int t = 0; t < nj; t++) pthread_create(&th[t], NULL, run_job, &jobs[t]);
    for (int t = 0; t < nj; t++) pthread_join(th[t], NULL);
    free(th);
  }
  /* concatenate in order; fix up diag line numbers */
  obuf tsv = {0}, loci = {0}, dos = {0};
  diag_sink ds = {0};
  uint64_t line_base = 0;
  for (int t = 0; t < nj; t++) {
    worker *w = &jobs[t].w;
    if (nj == 1) {
      tsv = w->tsv;
      loci = w->loci;
      dos = w->dosage;
    } else {
      ob_put(&tsv, w->tsv.p, w->tsv.n);
      ob_put(&loci, w->loci.p, w->loci.n);
      ob_put(&dos, w->dosage.p, w->dosage.n);
      free(w->tsv.p);
      free(w->loci.p);
      free(w->dosage.p);
    }
    res->n_rows += w->n_rows;
    res->n_dosage_rows += w->n_dosage_rows;
    for (size_t i = 0; i < w->ds.n; i++) {
      ds.line_no = w->ds.d[i].line_no + line_base;
      diag(&ds, w->ds.d[i].alt_no, w->ds.d[i].code);
    }
    free(w->ds.d);
    line_base += jobs[t].n_lines;
  }
  res->n_lines = line_base;
  res->tsv = tsv.p;
  res->tsv_len = tsv.n;
  res->loci = loci.p;
  res->loci_len = loci.n;
  res->dosage = (int8_t *)dos.p;
  res->diags = ds.d;
  res->n_diags = ds.n;
  res->n_samples = nh > 9 ? (uint32_t)(nh - 9) : 0;
  free(jobs);
  return 0;
}

int oracle_process_block(const oracle_config *cfg, const char *chrom_line, size_t chrom_len, int eol_width,
                         const char *block, size_t len, int threads, oracle_result *res) {
  memset(res, 0, sizeof(*res));
  sv *hdr;
  int nh;
  char *owned;
  /* chomp the header line if the caller left its EOL on */
  while (chrom_len && (chrom_line[chrom_len - 1] == '\n' || chrom_line[chrom_len - 1] == '\r')) chrom_len--;
  parse_header_line(cfg, chrom_line, chrom_len, &hdr, &nh, &owned);
  int rc = run_block(cfg, hdr, nh, eol_width, '\n', block, len, threads, res);
  free(hdr);
  free(owned);
  return rc;
}

/* readVcf main.go:241-396 */
int oracle_read_vcf(const oracle_config *cfg, const char *in, size_t len, int threads, oracle_result *res) {
  memset(res, 0, sizeof(*res));
  /* parse.FindEndOfLine :250 -- first line up to \r or \n; "\r\n" => numChars 2 */
  size_t i = 0;
  while (i < len && in[i] != '\n' && in[i] != '\r') i++;
  if (i >= len) {
    res->error = 1;
    return 1;
  }
  char eol = '\n';
  int eol_width = 1;
  size_t p = i + 1;
  if (in[i] == '\r') {
    if (p < len && in[p] == '\n') {
      eol_width = 2;
      p++;
    } else {
      eol = '\r';
    }
  }
  /* regexp "##fileformat=VCFv4" :256 */
  if (!memmem(in, i, "##fileformat=VCFv4", 18)) {
    res->error = 1;
    return 1;
  }
  /* find "#CHROM" :266-290 */
  sv *hdr = NULL;
  int nh = 0;
  char *owned = NULL;
  int found = 0;
  while (p < len) {
    const char *nl = (const char *)memchr(in + p, eol, len - p);
    if (!nl) break; /* io.EOF */
    size_t e = (size_t)(nl - in) + 1;
    size_t rl = e - p;
    size_t cl = rl >= (size_t)eol_width ? rl - (size_t)eol_width : 0;
    const char *tab = (const char *)memchr(in + p, '\t', cl);
    size_t f0 = tab ? (size_t)(tab - (in + p)) : cl;
    if (f0 == 6 && memcmp(in + p, "#CHROM", 6) == 0) {
      parse_header_line(cfg, in + p, cl, &hdr, &nh, &owned);
      found = 1;
      p = e;
      break;
    }
    p = e;
  }
  if (!found) {
    res->error = 2;
    return 2;
  }
  int rc = run_block(cfg, hdr, nh, eol_width, eol, in + p, len - p, threads, res);
  free(hdr);
  free(owned);
  return rc;
}

void oracle_free_result(oracle_result *res) {
  free(res->tsv);
  free(res->loci);
  free(res->dosage);
  free(res->diags);
  memset(res, 0, sizeof(*res));
}
