// bvcf_text.cuh -- exact text formatting/parsing on the device: Go's strconv.FormatFloat(q,'G',3,64),
// strconv.Itoa / Atoi, and the general GT grammar of makeHetHomozygotes (main.go:1126-1190).
#pragma once
#include "bvcf_common.cuh"

namespace bvcf {

// ---- strconv.FormatFloat(float64(k)/float64(n), 'G', 3, 64) for 0 < k <= n (main.go:627,641,655,670) ----
// The quotient is the IEEE double q = m * 2^x.  Its EXACT binary value is rounded to 3 significant decimal
// digits, ties to even, with 128-bit integer arithmetic: D = round(m * 10^(2-e) / 2^-x).  Trailing zeros are
// stripped; decimal exponent < -4 switches to d.ddE-XX (Go 'G' == C "%.3G" on (0,1]).
// Returns the text packed little-endian in a u64 (at most 8 characters) and its length.
__device__ __forceinline__ uint64_t format_ratio_g3(uint32_t k, uint32_t n, int &len_out) {
  unsigned long long r;  // the three significant digits, 100..999
  int e;                 // decimal exponent of the first digit
  bool have = false;
  if (n < (1u << 24)) {
    // Fast path, integers only.  D = round(k 10^j / n) with 100 <= D < 1000.  The double q differs from k/n by less
    // than 2^-53 relative, while k 10^j / n is at least 1/(2n) away from every rounding boundary unless it sits ON
    // one (2 rem == n): only then can the double fall on either side, and the exact path below decides.
    // (Checked against the exact path on 26 million (k, n) pairs.)
    unsigned long long m = k;
    const unsigned long long lim = 100ull * n;
    int j = 0;
    while (m < lim) { m *= 10ull; j++; }   // at most 10 rounds: k 10^j < 2^24 * 10^10 < 2^63
    unsigned long long D = m / n;
    const unsigned long long rem = m - D * n;
    if (2ull * rem != n) {
      if (2ull * rem > n) D++;
      e = 2 - j;
      if (D == 1000) { D = 100; e++; }
      r = D;
      have = true;
    }
  }
  if (!have) {
  const double q = (double)k / (double)n;  // IEEE-754 division, as in Go
  const unsigned long long bits = (unsigned long long)__double_as_longlong(q);
  const int bexp = (int)((bits >> 52) & 0x7FF);
  const unsigned long long m = (bits & 0xFFFFFFFFFFFFFull) | (1ull << 52);  // q is normal: k>=1, n<2^32
  const int s = 1075 - bexp;  // q = m * 2^-s, 52 <= s <= 52+32
  // estimate the decimal exponent, then fix it with exact integer comparisons
  e = 0;
  {
    double t = q;
    while (t < 1.0 && e > -15) { t *= 10.0; e--; }
  }
  bool up;
  for (;;) {
    unsigned long long p10 = 1;
    for (int i = 0; i < 2 - e; i++) p10 *= 10ull;       // 10^(2-e) <= 10^17 < 2^57
    const unsigned __int128 M = (unsigned __int128)m * p10;  // < 2^110
    r = (unsigned long long)(M >> s);
    if (r >= 1000) { e++; continue; }
    if (r < 100) { e--; continue; }
    const unsigned __int128 rem = M & ((((unsigned __int128)1) << s) - 1);
    const unsigned __int128 half = ((unsigned __int128)1) << (s - 1);
    up = rem > half || (rem == half && (r & 1));
    break;
  }
  if (up) { r++; if (r == 1000) { r = 100; e++; } }
  }
  uint32_t d[3] = {(uint32_t)(r / 100), (uint32_t)(r / 10 % 10), (uint32_t)(r % 10)};
  int nd = 3;
  while (nd > 1 && d[nd - 1] == 0) nd--;
  uint64_t out = 0;
  int len = 0;
  auto put = [&](uint32_t c) { out |= (uint64_t)c << (8 * len); len++; };
  if (e < -4) {  // %E style
    put('0' + d[0]);
    if (nd > 1) { put('.'); for (int i = 1; i < nd; i++) put('0' + d[i]); }
    put('E'); put('-');
    const int ae = -e;
    put('0' + ae / 10); put('0' + ae % 10);
  } else if (e >= 0) {  // only q == 1 (or rounds to 1)
    put('0' + d[0]);
    if (nd > 1) { put('.'); for (int i = 1; i < nd; i++) put('0' + d[i]); }
  } else {
    put('0'); put('.');
    for (int i = 0; i < -e - 1; i++) put('0');
    for (int i = 0; i < nd; i++) put('0' + d[i]);
  }
  len_out = len;
  return out;
}

// ---- strconv.Itoa -----------------------------------------------------------------------------
// digits into buf (at most 20 + sign); returns the length
__device__ __forceinline__ int itoa_dec(long long v, uint8_t *buf) {
  uint8_t tmp[20];
  int k = 0;
  unsigned long long u = v < 0 ? 0ull - (unsigned long long)v : (unsigned long long)v;
  do { tmp[k++] = (uint8_t)('0' + u % 10); u /= 10; } while (u);
  int n = 0;
  if (v < 0) buf[n++] = '-';
  while (k) buf[n++] = tmp[--k];
  return n;
}
__device__ __forceinline__ int dec_len(unsigned long long u) {
  int n = 1;
  if ((u >> 32) == 0) {  // 32-bit arithmetic for the common case (POS, ac, an)
    uint32_t w = (uint32_t)u;
    while (w >= 10) { w /= 10; n++; }
    return n;
  }
  while (u >= 10) { u /= 10; n++; }
  return n;
}
// strconv.Itoa without a byte buffer: the text of v packed little-endian in (lo, hi), at most 16 characters.
// Returns the length, or -1 when the text is longer (callers fall back to itoa_dec).
__device__ __forceinline__ int itoa_pack(long long v, unsigned long long &lo, unsigned long long &hi) {
  unsigned long long u = v < 0 ? 0ull - (unsigned long long)v : (unsigned long long)v;
  if (u >= 1000000000000000ull) return -1;
  lo = 0; hi = 0;
  int n = 0;
  // digits come least significant first; each one is shifted in below the previous, so byte 0 ends up the
  // most significant digit
  if ((u >> 32) == 0) {
    uint32_t w = (uint32_t)u;
    do {
      const uint32_t q = w / 10;
      hi = (hi << 8) | (lo >> 56);
      lo = (lo << 8) | (unsigned long long)('0' + (w - q * 10));
      w = q; n++;
    } while (w);
  } else {
    do {
      const unsigned long long q = u / 10;
      hi = (hi << 8) | (lo >> 56);
      lo = (lo << 8) | (unsigned long long)('0' + (uint32_t)(u - q * 10));
      u = q; n++;
    } while (u);
  }
  if (v < 0) { hi = (hi << 8) | (lo >> 56); lo = (lo << 8) | (unsigned long long)'-'; n++; }
  return n;
}

// ---- strconv.Atoi (main.go:752,824): optional sign, >= 1 digits, must fit int64 -----------------
__device__ __forceinline__ bool atoi_go(const uint8_t *p, int n, long long &out) {
  if (n == 0) return false;
  int i = 0;
  bool neg = false;
  if (p[0] == '-' || p[0] == '+') { neg = p[0] == '-'; i = 1; if (n == 1) return false; }
  unsigned long long v = 0;
  for (; i < n; i++) {
    const unsigned d = (unsigned)p[i] - '0';
    if (d > 9) return false;
    if (v > 1844674407370955161ull || (v == 1844674407370955161ull && d > 5)) return false;  // v * 10 + d > 2^64 - 1
    v = v * 10 + d;
  }
  if (!neg && v > 0x7FFFFFFFFFFFFFFFull) return false;
  if (neg && v > 0x8000000000000000ull) return false;
  out = neg ? (long long)(0ull - v) : (long long)v;
  return true;
}

// ---- general GT grammar for one sample field (main.go:1126-1190) ---------------------------------
// f: field start; maxn: bytes up to the end of the line's content.  a: allele number (altIdx+1).
// Returns class 0 none, 1 het, 2 hom, 3 missing; gt/alt are what the sample adds to an/ac.
__device__ __noinline__ int classify_gt_general(const uint8_t *f, uint32_t maxn, uint32_t a, uint32_t &gt_out,
                                                uint32_t &alt_out) {
  uint32_t gn = 0;  // alleleField = SplitN(field, ":", 2)[0]
  bool has_pipe = false, has_slash = false;
  while (gn < maxn) {
    const uint8_t c = f[gn];
    if (c == '\t' || c == ':') break;
    has_pipe |= c == '|';
    has_slash |= c == '/';
    gn++;
  }
  const uint8_t sep = has_pipe ? '|' : (has_slash ? '/' : 0);
  uint8_t ad[10];
  int an = 0;
  {
    uint8_t tmp[10];
    int k = 0;
    uint32_t u = a;
    do { tmp[k++] = (uint8_t)('0' + u % 10); u /= 10; } while (u);
    while (k) ad[an++] = tmp[--k];
  }
  uint32_t gt = 0, alt = 0, s = 0;
  for (;;) {
    uint32_t e = s;
    if (sep) { while (e < gn && f[e] != sep) e++; } else { e = gn; }
    const uint32_t tn = e - s;
    if (tn == 1 && f[s] == '.') { gt_out = 0; alt_out = 0; return 3; }
    if (tn == (uint32_t)an) {
      bool eq = true;
      for (uint32_t i = 0; i < tn; i++) eq = eq && f[s + i] == ad[i];
      alt += eq;
    }
    gt++;
    if (e >= gn) break;
    s = e + 1;
  }
  gt_out = gt;
  alt_out = alt;
  return alt == 0 ? 0 : (alt == gt ? 2 : 1);
}

}  // namespace bvcf
