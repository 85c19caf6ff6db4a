"""The device formatter of FormatFloat(k/n, 'G', 3, 64) (bvcf_text.cuh format_ratio_g3) restated in C and checked on the
CPU against C's "%.3G" (== Go's 'G',3 on (0, 1]): the integer fast path and the exact 128-bit path
must both agree with it, and the fast path must only give up on exact rational ties."""
import os
import subprocess
import textwrap

SRC = textwrap.dedent(r"""
    #include <stdint.h>
    #include <stdio.h>
    #include <stdlib.h>
    #include <string.h>
    typedef unsigned __int128 u128;
    /* exact path: bvcf_text.cuh, the `!have` branch */
    static void exact(uint32_t k, uint32_t n, uint64_t *r_out, int *e_out) {
      const double q = (double)k / (double)n;
      uint64_t bits; memcpy(&bits, &q, 8);
      const int bexp = (int)((bits >> 52) & 0x7FF);
      const uint64_t m = (bits & 0xFFFFFFFFFFFFFull) | (1ull << 52);
      const int s = 1075 - bexp;
      int e = 0;
      { double t = q; while (t < 1.0 && e > -15) { t *= 10.0; e--; } }
      uint64_t r; int up;
      for (;;) {
        uint64_t p10 = 1; for (int i = 0; i < 2 - e; i++) p10 *= 10ull;
        const u128 M = (u128)m * p10;
        r = (uint64_t)(M >> s);
        if (r >= 1000) { e++; continue; }
        if (r < 100) { e--; continue; }
        const u128 rem = M & ((((u128)1) << s) - 1), half = ((u128)1) << (s - 1);
        up = rem > half || (rem == half && (r & 1));
        break;
      }
      if (up) { r++; if (r == 1000) { r = 100; e++; } }
      *r_out = r; *e_out = e;
    }
    /* fast path: bvcf_text.cuh, the `n < 2^24` branch; 0 = falls back.  j and D are guessed in single precision and
       made exact with integer comparisons; `dj`, `dd` perturb the guesses the way a sloppy divide (__fdividef) could */
    static const uint64_t P10[20] = {1ull, 10ull, 100ull, 1000ull, 10000ull, 100000ull, 1000000ull, 10000000ull,
        100000000ull, 1000000000ull, 10000000000ull, 100000000000ull, 1000000000000ull, 10000000000000ull,
        100000000000000ull, 1000000000000000ull, 10000000000000000ull, 100000000000000000ull, 1000000000000000000ull,
        10000000000000000000ull};
    static int fast(uint32_t k, uint32_t n, uint64_t *r_out, int *e_out, int dj, int dd) {
      if (n >= (1u << 24)) return 0;
      const float qf = (float)k / (float)n;
      int j = 2 + (qf < 1.0f) + (qf < 0.1f) + (qf < 0.01f) + (qf < 0.001f) + (qf < 1e-4f) + (qf < 1e-5f) + (qf < 1e-6f) + (qf < 1e-7f);
      j += dj; if (j < 0) j = 0; if (j > 10) j = 10;
      uint64_t m = (uint64_t)k * P10[j];
      const uint64_t lim = 100ull * n;
      while (m < lim) { m *= 10ull; j++; }
      while (m >= 10ull * lim) { j--; m = (uint64_t)k * P10[j]; }
      uint64_t D = (uint64_t)((float)m / (float)n);
      D = (int64_t)D + dd < 0 ? 0 : D + dd;
      if (D > 999ull) D = 999ull;
      long long rem = (long long)m - (long long)(D * n);
      while (rem < 0) { D--; rem += n; }
      while (rem >= (long long)n) { D++; rem -= n; }
      if (2ull * (uint64_t)rem == n) return 0;
      if (2ull * (uint64_t)rem > n) D++;
      int e = 2 - j;
      if (D == 1000) { D = 100; e++; }
      *r_out = D; *e_out = e; return 1;
    }
    /* digits + exponent -> text, as the device does (g3_text: packed little-endian in a u64) */
    static void text(uint64_t r64, int e, char *outc) {
      const uint32_t r = (uint32_t)r64;
      const uint32_t d0 = r / 100u, r2 = r - d0 * 100u, d1 = r2 / 10u, d2 = r2 - d1 * 10u;
      const int nd = d2 ? 3 : (d1 ? 2 : 1);
      uint64_t out; int len;
      if (e < -4 || e >= 0) {
        out = '0' + d0; len = 1;
        if (nd > 1) {
          out |= (uint64_t)'.' << 8 | (uint64_t)('0' + d1) << 16; len = 3;
          if (nd > 2) { out |= (uint64_t)('0' + d2) << 24; len = 4; }
        }
        if (e < -4) {
          const uint32_t ae = (uint32_t)(-e), t = ae / 10u;
          out |= ((uint64_t)'E' | (uint64_t)'-' << 8 | (uint64_t)('0' + t) << 16 | (uint64_t)('0' + (ae - t * 10u)) << 24) << (8 * len);
          len += 4;
        }
      } else {
        const int z = -e - 1;
        const uint64_t digits = (uint64_t)('0' + d0) | (uint64_t)('0' + d1) << 8 | (uint64_t)('0' + d2) << 16;
        const uint64_t keep = nd == 3 ? 0xFFFFFFull : (nd == 2 ? 0xFFFFull : 0xFFull);
        const uint64_t pre = 0x3030303030302E30ull & ((1ull << (8 * (2 + z))) - 1ull);
        len = 2 + z + nd;
        out = pre | (digits & keep) << (8 * (2 + z));
      }
      for (int i = 0; i < len; i++) outc[i] = (char)(out >> (8 * i));
      outc[len] = 0;
    }
    int main(void) {
      const uint32_t ns[] = {5008, 2504, 400000, 399996, 3, 6, 7, 10, 16, 1000, 4096, 65535, 1u << 20, 16777215, 16777216, 50000000};
      long bad = 0, fb = 0, ties = 0, n_checked = 0;
      for (unsigned i = 0; i < sizeof(ns) / sizeof(ns[0]); i++) {
        const uint32_t n = ns[i], step = n > 2000000 ? 9973 : (n > 100000 ? 7 : 1);
        for (uint32_t k = 1; k <= n; k += step) {
          char want[32], a[32], b[32];
          snprintf(want, sizeof want, "%.3G", (double)k / (double)n);
          uint64_t r; int e;
          exact(k, n, &r, &e); text(r, e, a);
          if (strcmp(a, want)) { if (bad++ < 5) printf("exact %u/%u: %s != %s\n", k, n, a, want); }
          if (fast(k, n, &r, &e, 0, 0)) { text(r, e, b); if (strcmp(b, want)) { if (bad++ < 5) printf("fast %u/%u: %s != %s\n", k, n, b, want); } }
          else { fb++; if (n < (1u << 24)) ties++; }
          /* the guesses may be off: the integer corrections must make up for it */
          static const int pert[4][2] = {{-1, -3}, {1, 3}, {-2, 2}, {2, -2}};
          for (int pi = 0; pi < 4; pi++) {
            uint64_t r2; int e2;
            if (fast(k, n, &r2, &e2, pert[pi][0], pert[pi][1])) { text(r2, e2, b); if (strcmp(b, want)) { if (bad++ < 5) printf("fast' %u/%u: %s != %s\n", k, n, b, want); } }
          }
          n_checked++;
        }
      }
      printf("checked %ld fallbacks %ld ties %ld bad %ld\n", n_checked, fb, ties, bad);
      return bad != 0;
    }
""")


def test_device_float_formatter_restated_in_c(tmp_path):
    src = tmp_path / "fmt.c"
    exe = tmp_path / "fmt"
    src.write_text(SRC)
    subprocess.check_call(["gcc", "-O2", "-o", str(exe), str(src)])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout[-2000:]
    assert " bad 0" in out.stdout
    # below 2^24 the fast path only gives up on exact rational ties: a fraction of a percent
    tok = out.stdout.split()
    checked, ties = int(tok[tok.index("checked") + 1]), int(tok[tok.index("ties") + 1])
    assert ties < checked // 200


def test_c_restatement_matches_device_source():
    """the constants the proof rests on are the ones in the CUDA source"""
    here = os.path.dirname(os.path.abspath(__file__))
    cu = open(os.path.join(here, "..", "bystro_vcf_b200", "csrc", "bvcf_text.cuh")).read()
    for needle in ("if (n < (1u << 24))", "const unsigned long long lim = 100ull * n;", "if (2ull * (unsigned long long)rem != n)",
                   "if (2ull * (unsigned long long)rem > n) D++;", "if (D == 1000) { D = 100; e++; }",
                   "while (rem < 0) { D--; rem += n; }", "while (rem >= (long long)n) { D++; rem -= n; }",
                   "0x3030303030302E30ull"):
        assert needle in cu, needle
