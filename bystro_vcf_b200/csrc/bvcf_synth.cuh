// bvcf_synth.cuh -- seeded synthetic VCF workloads, identical on host and device.
//
// Every data line is a pure function of (seed, line number): counter-based hashing (SplitMix64), so any
// shard can be generated independently, on the host (CPU baseline sample, tests) or on the device (the
// full-size resident workloads never touch the host).  Bench/test infrastructure, not part of the transform.
//
// shapes: 0 = C2 chr1-shape (1000G Phase 3: phased diploid GT, AC spectrum of the reference's chr1 fixture)
//         1 = C3 sites-only (8 columns; 30 % multiallelic/MNP, padded indels, FILTER mix, CHROM cycling)
//         2 = C4 biobank (as C2 + 2 % missing genotypes)
//         3 = C5 (as C2 with the FILTER column re-drawn: 85 % PASS, 10 % LowQual, 5 % '.')
#pragma once
#include <cstdint>

#ifdef __CUDACC__
#define SYN_HD __host__ __device__ __forceinline__
#else
#define SYN_HD inline
#endif

namespace bvcf_synth {

struct Params {
  uint64_t seed;
  uint32_t n_samples;
  int shape;
};

struct LineGeno {
  uint32_t n_alts;       // ALT alleles on the line (1..6)
  uint32_t thr;          // haplotype is ALT when hash32 < thr
  uint32_t forced_hap;   // this haplotype is always ALT (every line has ac >= 1)
  uint32_t miss_thr;     // sample is ".|." when hash32 < miss_thr
};

SYN_HD uint64_t mix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
SYN_HD uint64_t h3(uint64_t seed, uint64_t line, uint64_t k) { return mix64(mix64(seed ^ (line * 0xD1B54A32D192ED03ull)) + k); }

struct Buf {
  uint8_t *p;
  uint32_t n;
  SYN_HD void c(uint8_t ch) { p[n++] = ch; }
  SYN_HD void s(const char *t) { while (*t) p[n++] = (uint8_t)*t++; }
  SYN_HD void u(uint64_t v) {
    uint8_t tmp[20];
    int k = 0;
    do { tmp[k++] = (uint8_t)('0' + v % 10); v /= 10; } while (v);
    while (k) p[n++] = tmp[--k];
  }
  SYN_HD void frac(uint32_t v, int digits) {  // "0.dddd"
    p[n++] = '0'; p[n++] = '.';
    uint32_t d = 1;
    for (int i = 1; i < digits; i++) d *= 10;
    for (int i = 0; i < digits; i++) { p[n++] = (uint8_t)('0' + (v / d) % 10); d = d > 1 ? d / 10 : 1; }
  }
};

SYN_HD uint8_t base_of(uint64_t h) { return (uint8_t)"ACGT"[h & 3]; }

// target alt-allele count: the reference fixture's spectrum, log-uniform inside a bin
SYN_HD uint32_t draw_ac(uint64_t h, uint32_t n_hap) {
  const uint32_t u = (uint32_t)(h % 1000);
  const uint32_t r = (uint32_t)((h >> 20) & 0xFFFFF);  // 20 random bits
  uint32_t lo, hi;
  if (u < 386) return 1;
  if (u < 501) return 2;
  if (u < 706) { lo = 3; hi = 10; }
  else if (u < 878) { lo = 11; hi = 100; }
  else if (u < 949) { lo = 101; hi = 1000; }
  else { lo = 1001; hi = 4900; }
  // log-uniform: lo * (hi/lo)^(r/2^20), integer approximation by repeated square-root-free stepping
  double f = (double)r / 1048576.0;
  double v = (double)lo;
  double ratio = (double)hi / (double)lo;
  // v = lo * ratio^f via 20 binary digits of f
  double rt = ratio;
  for (int i = 0; i < 12; i++) {
    // rt = sqrt(rt) by Newton (deterministic on host and device: only + * /)
    double x = rt > 1.0 ? rt : 1.0;
    for (int k = 0; k < 6; k++) x = 0.5 * (x + rt / x);
    rt = x;
    f *= 2.0;
    if (f >= 1.0) { v *= rt; f -= 1.0; }
  }
  uint32_t ac = (uint32_t)v;
  if (ac < lo) ac = lo;
  if (ac > hi) ac = hi;
  // scale the spectrum to the haplotype count (5008 in the fixture)
  const uint64_t scaled = (uint64_t)ac * n_hap / 5008ull;
  return scaled < 1 ? 1u : (uint32_t)scaled;
}

// Writes CHROM..FORMAT (with the trailing tab when samples follow, or the '\n' when there are none) into
// buf (>= 768 bytes) and returns the length; fills g.
SYN_HD uint32_t line_prefix(const Params &P, uint64_t line, uint8_t *out, LineGeno &g) {
  Buf b{out, 0};
  const uint64_t hA = h3(P.seed, line, 1), hB = h3(P.seed, line, 2), hC = h3(P.seed, line, 3), hD = h3(P.seed, line, 4);
  const bool sites = P.shape == 1;
  // CHROM
  if (!sites) {
    b.c('1');
  } else {
    const uint32_t c = (uint32_t)(line / 4096 % 27);
    if (c < 22) b.u(c + 1);
    else if (c == 22) b.c('X');
    else if (c == 23) b.c('Y');
    else if (c == 24) b.s("MT");
    else if (c == 25) b.s("chr1");
    else b.s("GL000207.1");
  }
  b.c('\t');
  b.u(10177ull + 81ull * line + hA % 81);  // POS strictly increasing
  b.c('\t');
  b.s("rs"); b.u(100000 + (hA >> 8) % 900000000ull);
  if ((hA >> 40) % 500 == 0) { b.s(";rs"); b.u(100000 + (hA >> 16) % 900000000ull); }
  b.c('\t');
  // REF / ALT
  uint32_t n_alts = 1;
  const uint32_t cls = (uint32_t)(hB % 10000);
  const uint8_t r0 = base_of(hB >> 16);
  uint8_t ref[40];
  uint32_t ref_n = 0;
  auto other = [&](uint8_t base, uint64_t h) { return base_of((uint64_t)(((base == 'A') ? 0 : (base == 'C') ? 1 : (base == 'G') ? 2 : 3) + 1 + h % 3)); };
  if (!sites) {
    if ((hB >> 32) % 1000 < 7) n_alts = 2 + (uint32_t)((hB >> 44) % 2);  // 0.65 % multi-ALT lines
    uint32_t kind = cls < 9590 ? 0 : cls < 9840 ? 1 : cls < 9980 ? 2 : cls < 9990 ? 3 : 4;
    const uint32_t k = 1 + (uint32_t)((hB >> 24) % 8);
    if (kind == 1) { ref[ref_n++] = r0; for (uint32_t i = 0; i < k; i++) ref[ref_n++] = base_of(hC >> (2 * i)); }
    else if (kind == 3) { for (uint32_t i = 0; i < 2 + k % 3; i++) ref[ref_n++] = base_of(hC >> (2 * i)); }
    else ref[ref_n++] = r0;
    for (uint32_t i = 0; i < ref_n; i++) b.c(ref[i]);
    b.c('\t');
    for (uint32_t a = 0; a < n_alts; a++) {
      if (a) b.c(',');
      const uint64_t ha = hD >> (7 * a);
      const uint32_t kk = a == 0 ? kind : (uint32_t)(ha % 3 == 0 ? 2 : 0);
      if (kk == 0) { b.c(other(ref[0], ha + a)); for (uint32_t i = 1; i < ref_n; i++) b.c(ref[i]); }
      else if (kk == 1) b.c(r0);
      else if (kk == 2) { for (uint32_t i = 0; i < ref_n; i++) b.c(ref[i]); for (uint32_t i = 0; i < 1 + (uint32_t)(ha % 6); i++) b.c(base_of(ha >> (2 * i + 3))); }
      else if (kk == 3) { for (uint32_t i = 0; i < ref_n; i++) b.c(((ha >> i) & 1) || i == 0 ? other(ref[i], ha >> (3 * i)) : ref[i]); }
      else b.s("<CN0>");
    }
  } else {
    // C3: parse/normalise-bound mix
    const uint32_t u = cls % 1000;
    const uint32_t pad_l = 1 + (uint32_t)((hB >> 24) % 6), pad_r = (uint32_t)((hB >> 28) % 5);
    uint32_t kind;  // 0 snp, 1 del, 2 ins, 3 mnp, 4 multi, 5 invalid, 6 mixed
    if (u < 5) kind = 5; else if (u < 15) kind = 6; else if (u < 165) kind = 4; else if (u < 315) kind = 3;
    else if (u < 415) kind = 1; else if (u < 515) kind = 2; else kind = 0;
    const uint32_t core = 1 + (uint32_t)((hB >> 34) % 7);
    if (kind == 0) { ref[ref_n++] = r0; }
    else if (kind == 3) { for (uint32_t i = 0; i < 2 + core % 7; i++) ref[ref_n++] = base_of(hC >> (2 * i)); }
    else if (kind == 1 || kind == 6) { for (uint32_t i = 0; i < pad_l + core + pad_r; i++) ref[ref_n++] = base_of(hC >> (2 * (i % 30))); }
    else if (kind == 2) { for (uint32_t i = 0; i < pad_l + pad_r; i++) ref[ref_n++] = base_of(hC >> (2 * i)); }
    else { ref[ref_n++] = r0; if (kind == 4 && (hB >> 40) % 2) { ref[ref_n++] = base_of(hC); ref[ref_n++] = base_of(hC >> 2); } }
    for (uint32_t i = 0; i < ref_n; i++) b.c(ref[i]);
    b.c('\t');
    if (kind == 0) b.c(other(r0, hD));
    else if (kind == 3) { for (uint32_t i = 0; i < ref_n; i++) b.c(((hD >> i) & 1) || i == ref_n - 1 ? other(ref[i], hD >> (3 * i)) : ref[i]); }
    else if (kind == 1) { for (uint32_t i = 0; i < pad_l; i++) b.c(ref[i]); for (uint32_t i = 0; i < pad_r; i++) b.c(ref[pad_l + core + i]); }
    else if (kind == 6) { b.c(other(ref[0], hD)); for (uint32_t i = 1; i < pad_l; i++) b.c(ref[i]); }
    else if (kind == 2) { for (uint32_t i = 0; i < pad_l; i++) b.c(ref[i]); for (uint32_t i = 0; i < core; i++) b.c(base_of(hD >> (2 * i))); for (uint32_t i = 0; i < pad_r; i++) b.c(ref[pad_l + i]); }
    else if (kind == 5) { const uint32_t w = (uint32_t)(hD % 5); b.s(w == 0 ? "." : w == 1 ? "*" : w == 2 ? "<DEL>" : w == 3 ? "a" : "N"); }
    else {
      n_alts = 2 + (uint32_t)(hD % 5);
      for (uint32_t a = 0; a < n_alts; a++) {
        if (a) b.c(',');
        const uint64_t ha = mix64(hD + a);
        const uint32_t w = (uint32_t)(ha % 4);
        if (w == 0 || ref_n == 1) { b.c(w == 3 ? ref[0] : other(ref[0], ha)); for (uint32_t i = 1; i < ref_n; i++) b.c(ref[i]); if (w == 3 || w == 2) for (uint32_t i = 0; i < 1 + (uint32_t)(ha >> 8) % 4; i++) b.c(base_of(ha >> (10 + 2 * i))); }
        else if (w == 1) b.c(ref[0]);
        else if (w == 2) { for (uint32_t i = 0; i < ref_n; i++) b.c(ref[i]); for (uint32_t i = 0; i < 1 + (uint32_t)(ha >> 8) % 4; i++) b.c(base_of(ha >> (10 + 2 * i))); }
        else { b.c(ref[0]); b.c(other(ref[1], ha)); for (uint32_t i = 2; i < ref_n; i++) b.c(ref[i]); }
      }
    }
  }
  b.c('\t');
  b.s("100");
  b.c('\t');
  // FILTER
  {
    const uint32_t f = (uint32_t)((hA >> 50) % 100);
    if (P.shape == 3) b.s(f < 85 ? "PASS" : f < 95 ? "LowQual" : ".");
    else if (sites) b.s(f < 90 ? "PASS" : f < 95 ? "." : f < 98 ? "q10" : "LowQual;q10");
    else b.s("PASS");
  }
  b.c('\t');
  // INFO (1000G-style, ~115 bytes)
  const uint32_t n_hap = 2 * (P.n_samples ? P.n_samples : 2504);
  const uint32_t ac = draw_ac(hC, n_hap);
  b.s("AC="); b.u(ac);
  b.s(";AF="); b.frac((uint32_t)((uint64_t)ac * 1000000ull / n_hap), 6);
  b.s(";AN="); b.u(n_hap);
  b.s(";NS="); b.u(n_hap / 2);
  b.s(";DP="); b.u(10000 + (hD >> 10) % 20000);
  b.s(";EAS_AF="); b.frac((uint32_t)((hD >> 20) % 10000), 4);
  b.s(";AMR_AF="); b.frac((uint32_t)((hD >> 30) % 10000), 4);
  b.s(";AFR_AF="); b.frac((uint32_t)((hD >> 40) % 10000), 4);
  b.s(";EUR_AF="); b.frac((uint32_t)((hC >> 40) % 10000), 4);
  b.s(";SAS_AF="); b.frac((uint32_t)((hC >> 50) % 10000), 4);
  b.s(";AA="); b.c(r0); b.s("|||;VT=SNP");
  if (P.n_samples == 0) {
    b.c('\n');
  } else {
    b.s("\tGT\t");
  }
  g.n_alts = n_alts > 3 ? 3 : n_alts;
  g.thr = ac <= 1 ? 0u : (uint32_t)(((uint64_t)(ac - 1) << 32) / n_hap > 0xFFFFFFFFull ? 0xFFFFFFFFull : ((uint64_t)(ac - 1) << 32) / n_hap);
  g.forced_hap = (uint32_t)(hD % n_hap);
  g.miss_thr = P.shape == 2 ? (uint32_t)(0.02 * 4294967296.0) : 0u;
  return b.n;
}

// j-th byte of the sample region of `line`: tokens are "a|b" + '\t' ('\n' after the last sample)
SYN_HD uint8_t gt_byte(const Params &P, const LineGeno &g, uint64_t line, uint64_t j) {
  const uint32_t s = (uint32_t)(j >> 2), k = (uint32_t)(j & 3);
  if (k == 1) return '|';
  if (k == 3) return s + 1 == P.n_samples ? '\n' : '\t';
  const uint64_t hs = h3(P.seed, line, 1000 + s);
  if (g.miss_thr && (uint32_t)(hs >> 32) < g.miss_thr) return '.';
  const uint32_t hap = 2 * s + (k >> 1);
  const uint32_t hv = k == 0 ? (uint32_t)hs : (uint32_t)(hs >> 16) ^ (uint32_t)(hs >> 33);
  const bool alt = hap == g.forced_hap || hv < g.thr;
  if (!alt) return '0';
  return (uint8_t)('1' + (g.n_alts > 1 ? (hs >> 50) % g.n_alts : 0));
}

SYN_HD uint64_t line_len(const Params &P, uint32_t prefix_len) { return (uint64_t)prefix_len + 4ull * P.n_samples; }

}  // namespace bvcf_synth
