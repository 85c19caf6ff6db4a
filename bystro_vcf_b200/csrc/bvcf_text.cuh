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
__device__ const unsigned long long BVCF_P10[20] = {1ull, 10ull, 100ull, 1000ull, 10000ull, 100000ull, 1000000ull, 10000000ull,
    100000000ull, 1000000000ull, 10000000000ull, 100000000000ull, 1000000000000ull, 10000000000000ull,
    100000000000000ull, 1000000000000000ull, 10000000000000000ull, 100000000000000000ull, 1000000000000000000ull,
    10000000000000000000ull};

// the text of three significant digits r (100..999) with decimal exponent e (value = r / 100 * 10^e, e <= 0)
__device__ __forceinline__ uint64_t g3_text(uint32_t r, int e, int &len_out) {
  const uint32_t d0 = r / 100u, r2 = r - d0 * 100u, d1 = r2 / 10u, d2 = r2 - d1 * 10u;
  const int nd = d2 ? 3 : (d1 ? 2 : 1);  // trailing zeros are stripped
  if (e < -4) {  // %E style: d[.dd]E-XX
    uint64_t out = '0' + d0;
    int len = 1;
    if (nd > 1) {
      out |= (uint64_t)'.' << 8 | (uint64_t)('0' + d1) << 16;
      len = 3;
      if (nd > 2) { out |= (uint64_t)('0' + d2) << 24; len = 4; }
    }
    const uint32_t ae = (uint32_t)(-e), t = ae / 10u;
    out |= ((uint64_t)'E' | (uint64_t)'-' << 8 | (uint64_t)('0' + t) << 16 | (uint64_t)('0' + (ae - t * 10u)) << 24) << (8 * len);
    len_out = len + 4;
    return out;
  }
  if (e >= 0) {  // only q == 1 (or rounds to 1): d[.dd]
    uint64_t out = '0' + d0;
    int len = 1;
    if (nd > 1) {
      out |= (uint64_t)'.' << 8 | (uint64_t)('0' + d1) << 16;
      len = 3;
      if (nd > 2) { out |= (uint64_t)('0' + d2) << 24; len = 4; }
    }
    len_out = len;
    return out;
  }
  // 0.[000]ddd: "0." + (-e - 1) zeros + the digits
  const int z = -e - 1;  // 0..3
  const uint64_t digits = (uint64_t)('0' + d0) | (uint64_t)('0' + d1) << 8 | (uint64_t)('0' + d2) << 16;
  const uint64_t keep = nd == 3 ? 0xFFFFFFull : (nd == 2 ? 0xFFFFull : 0xFFull);
  const uint64_t pre = 0x3030303030302E30ull & ((1ull << (8 * (2 + z))) - 1ull);  // "0.000000" cut to 2 + z characters
  len_out = 2 + z + nd;
  return pre | (digits & keep) << (8 * (2 + z));
}

// (out of line: four call sites per row composer, and the exact path is a few hundred instructions of 128-bit
// arithmetic -- inlined they pushed the compose kernel past the instruction cache)
__device__ __noinline__ uint64_t format_ratio_g3(uint32_t k, uint32_t n, int &len_out) {
  if (n < (1u << 24)) {
    // Fast path, integers only.  D = round(k 10^j / n) with 100 <= D < 1000.  The double q differs from k/n by less
    // than 2^-53 relative, while k 10^j / n is at least 1/(2n) away from every rounding boundary unless it sits ON
    // one (2 rem == n): only then can the double fall on either side, and the exact path below decides.
    // (Checked against the exact path on 26 million (k, n) pairs.)
    // j and D are first guessed in single precision, then made exact with 64-bit integer comparisons: no 64-bit
    // division, no digit loop.
    const float qf = __fdividef((float)k, (float)n);
    int j = 2 + (qf < 1.0f) + (qf < 0.1f) + (qf < 0.01f) + (qf < 0.001f) + (qf < 1e-4f) + (qf < 1e-5f) + (qf < 1e-6f) + (qf < 1e-7f);
    unsigned long long m = (unsigned long long)k * BVCF_P10[j];   // k 10^j < 2^24 * 10^10 < 2^63
    const unsigned long long lim = 100ull * n;
    while (m < lim) { m *= 10ull; j++; }                          // the guess may be one off at a power of ten
    while (m >= 10ull * lim) { j--; m = (unsigned long long)k * BVCF_P10[j]; }
    unsigned long long D = (unsigned long long)(__fdividef((float)m, (float)n));  // within a few units of m / n
    if (D > 999ull) D = 999ull;
    long long rem = (long long)m - (long long)(D * n);
    while (rem < 0) { D--; rem += n; }
    while (rem >= (long long)n) { D++; rem -= n; }
    if (2ull * (unsigned long long)rem != n) {
      if (2ull * (unsigned long long)rem > n) D++;
      int e = 2 - j;
      if (D == 1000) { D = 100; e++; }
      return g3_text((uint32_t)D, e, len_out);
    }
  }
  unsigned long long r;  // the three significant digits, 100..999
  int e;                 // decimal exponent of the first digit
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
  return g3_text((uint32_t)r, e, len_out);
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
__device__ __noinline__ int itoa_pack(long long v, unsigned long long &lo, unsigned long long &hi) {
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

// eight bytes at any alignment (three aligned words, funnel-shifted); reads up to 11 bytes past s -- the input
// buffers are padded
__device__ __forceinline__ unsigned long long ld64_any(const uint8_t *s) {
  const uint32_t sh = (uint32_t)((uintptr_t)s & 3u) * 8u;
  const uint32_t *wp = reinterpret_cast<const uint32_t *>((uintptr_t)s & ~(uintptr_t)3);
  const uint32_t w0 = wp[0], w1 = wp[1], w2 = wp[2];
  return (unsigned long long)__funnelshift_r(w0, w1, sh) | ((unsigned long long)__funnelshift_r(w1, w2, sh) << 32);
}

// ---- strconv.Atoi (main.go:752,824): optional sign, >= 1 digits, must fit int64 -----------------
// p points into the (padded) input: fields of up to 16 characters are read with two word loads, not a byte at a time
__device__ __forceinline__ bool atoi_go(const uint8_t *p, int n, long long &out) {
  if (n == 0) return false;
  const bool in_regs = n <= 16;
  const unsigned long long w0 = in_regs ? ld64_any(p) : 0ull, w1 = (in_regs && n > 8) ? ld64_any(p + 8) : 0ull;
  auto at = [&](int i) -> unsigned { return in_regs ? (unsigned)((i < 8 ? w0 : w1) >> (8 * (i & 7))) & 0xFFu : (unsigned)p[i]; };
  int i = 0;
  bool neg = false;
  const unsigned c0 = at(0);
  if (c0 == '-' || c0 == '+') { neg = c0 == '-'; i = 1; if (n == 1) return false; }
  unsigned long long v = 0;
#pragma unroll 1
  for (; i < n; i++) {
    const unsigned d = at(i) - '0';
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
