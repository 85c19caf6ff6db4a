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
    /* fast path: bvcf_text.cuh, the `n < 2^24` branch; 0 = falls back */
    static int fast(uint32_t k, uint32_t n, uint64_t *r_out, int *e_out) {
      if (n >= (1u << 24)) return 0;
      uint64_t m = k; int j = 0; const uint64_t lim = 100ull * n;
      while (m < lim) { m *= 10; j++; }
      uint64_t D = m / n, rem = m - D * n;
      if (2 * rem == n) return 0;
      if (2 * rem > n) D++;
      int e = 2 - j;
      if (D == 1000) { D = 100; e++; }
      *r_out = D; *e_out = e; return 1;
    }
    /* digits + exponent -> text, as the device does */
    static void text(uint64_t r, int e, char *out) {
      int d[3] = {(int)(r / 100), (int)(r / 10 % 10), (int)(r % 10)}, nd = 3, len = 0;
      while (nd > 1 && d[nd - 1] == 0) nd--;
      if (e < -4) {
        out[len++] = '0' + d[0];
        if (nd > 1) { out[len++] = '.'; for (int i = 1; i < nd; i++) out[len++] = '0' + d[i]; }
        out[len++] = 'E'; out[len++] = '-'; out[len++] = '0' + (-e) / 10; out[len++] = '0' + (-e) % 10;
      } else if (e >= 0) {
        out[len++] = '0' + d[0];
        if (nd > 1) { out[len++] = '.'; for (int i = 1; i < nd; i++) out[len++] = '0' + d[i]; }
      } else {
        out[len++] = '0'; out[len++] = '.';
        for (int i = 0; i < -e - 1; i++) out[len++] = '0';
        for (int i = 0; i < nd; i++) out[len++] = '0' + d[i];
      }
      out[len] = 0;
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
          if (fast(k, n, &r, &e)) { text(r, e, b); if (strcmp(b, want)) { if (bad++ < 5) printf("fast %u/%u: %s != %s\n", k, n, b, want); } }
          else { fb++; if (n < (1u << 24)) ties++; }
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
    for needle in ("if (n < (1u << 24))", "const unsigned long long lim = 100ull * n;", "if (2ull * rem != n)", "if (2ull * rem > n) D++;",
                   "if (D == 1000) { D = 100; e++; }"):
        assert needle in cu, needle
