// bvcf_inflate.cuh -- SURVEY 8f-3: bgzf / DEFLATE (RFC 1951) decompression on the GPU, so that the COMPRESSED bytes
// cross PCIe (a 1000 Genomes GT block deflates some 30x) and the uncompressed VCF text only ever exists in HBM,
// where the scan kernel reads it.  Upstream of main.go:192 the reference relies on `pigz -d -c` (README.md:10,46:
// "runs at pigz -p 1 limit").
//
// bgzf (the block gzip of bgzip / htslib, what .vcf.gz files are) cuts the stream into independent gzip members of
// at most 64 KiB of text, each with its compressed size in the header: the host walks the headers (a few bytes per
// block), the device inflates the blocks in parallel, ONE THREAD PER BLOCK.  That is the natural grain: inside a
// block DEFLATE is serial (Huffman codes of unknown length, matches that copy from the bytes just written).  A thread
// keeps a 64-bit bit buffer, canonical Huffman tables in its local memory (count / symbol arrays, decoded code length
// by code length as in zlib's puff.c -- small enough to stay in L1), and writes its text byte by byte; genotype text
// is long matches at distance 4 ("0|0\t" repeated), which is the tight inner loop.  Stored, fixed and dynamic blocks
// are all handled; the gzip CRC32 is not verified (ISIZE is).
#pragma once
#include "bvcf_common.cuh"

namespace bvcf {

struct InflateBlock {
  unsigned long long in_off;   // first byte of the block's DEFLATE payload in `comp`
  unsigned long long out_off;  // where its text goes in `out`
  uint32_t in_len;             // payload bytes
  uint32_t out_len;            // ISIZE: bytes of text
};
struct InflateParams {
  const uint8_t *comp;
  uint8_t *out;
  const InflateBlock *blocks;
  uint32_t n_blocks;
  uint32_t *n_bad;             // blocks that failed (corrupt stream, size mismatch)
};

constexpr int INF_MAXBITS = 15, INF_MAXL = 288, INF_MAXD = 30;

struct InfBits {
  const uint8_t *p;            // next input byte
  const uint8_t *end;
  unsigned long long buf;      // bits not yet consumed, LSB first
  int cnt;
  __device__ __forceinline__ void refill() {
    while (cnt <= 56) {
      const unsigned long long b = p < end ? (unsigned long long)*p : 0ull;  // zeros past the end: the size check catches a short stream
      p++;
      buf |= b << cnt;
      cnt += 8;
    }
  }
  __device__ __forceinline__ uint32_t bits(int n) {  // n <= 32
    if (cnt < n) refill();
    const uint32_t v = (uint32_t)(buf & ((1ull << n) - 1ull));
    buf >>= n; cnt -= n;
    return v;
  }
};

struct InfHuff {
  short *count;   // [INF_MAXBITS + 1] codes of each length
  short *symbol;  // symbols ordered by code
};

// canonical Huffman decode, one bit at a time (puff.c); at least 15 bits are in the buffer
__device__ __forceinline__ int inf_decode(InfBits &b, const InfHuff &h) {
  if (b.cnt < INF_MAXBITS) b.refill();
  int code = 0, first = 0, index = 0;
  unsigned long long buf = b.buf;
  for (int len = 1; len <= INF_MAXBITS; len++) {
    code |= (int)(buf & 1ull);
    buf >>= 1;
    const int count = h.count[len];
    if (code - count < first) {
      b.buf = buf; b.cnt -= len;
      return h.symbol[index + (code - first)];
    }
    index += count;
    first += count;
    first <<= 1;
    code <<= 1;
  }
  return -1;
}

// count / symbol tables from code lengths; returns < 0 for an over-subscribed set
__device__ __forceinline__ int inf_construct(InfHuff &h, const short *length, int n) {
  for (int len = 0; len <= INF_MAXBITS; len++) h.count[len] = 0;
  for (int s = 0; s < n; s++) h.count[length[s]]++;
  if (h.count[0] == n) return 0;
  int left = 1;
  for (int len = 1; len <= INF_MAXBITS; len++) {
    left <<= 1;
    left -= h.count[len];
    if (left < 0) return left;
  }
  short offs[INF_MAXBITS + 1];
  offs[1] = 0;
  for (int len = 1; len < INF_MAXBITS; len++) offs[len + 1] = offs[len] + h.count[len];
  for (int s = 0; s < n; s++)
    if (length[s] != 0) h.symbol[offs[length[s]]++] = (short)s;
  return left;
}

__device__ const short INF_LBASE[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
__device__ const short INF_LEXT[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
__device__ const short INF_DBASE[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
__device__ const short INF_DEXT[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
__device__ const unsigned char INF_ORDER[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

// literal/length + distance codes until the end-of-block symbol; returns 0 or an error
__device__ __forceinline__ int inf_codes(InfBits &b, const InfHuff &lc, const InfHuff &dc, uint8_t *out, uint32_t &o, uint32_t out_len) {
  for (;;) {
    int sym = inf_decode(b, lc);
    if (sym < 0) return 1;
    if (sym < 256) {
      if (o >= out_len) return 2;
      out[o++] = (uint8_t)sym;
    } else if (sym == 256) {
      return 0;
    } else {
      sym -= 257;
      if (sym >= 29) return 3;
      const uint32_t len = (uint32_t)INF_LBASE[sym] + b.bits(INF_LEXT[sym]);
      const int ds = inf_decode(b, dc);
      if (ds < 0 || ds >= 30) return 4;
      const uint32_t dist = (uint32_t)INF_DBASE[ds] + b.bits(INF_DEXT[ds]);
      if (dist > o) return 5;           // bgzf blocks are independent: no history before the block
      if (o + len > out_len) return 2;
      const uint8_t *src = out + o - dist;
      uint8_t *dst = out + o;
      for (uint32_t i = 0; i < len; i++) dst[i] = src[i];  // overlapping on purpose (dist < len repeats the pattern)
      o += len;
    }
  }
}

__global__ void __launch_bounds__(64) bvcf_inflate_kernel(const InflateParams p) {
  const uint32_t bi = blockIdx.x * blockDim.x + threadIdx.x;
  if (bi >= p.n_blocks) return;
  const InflateBlock blk = p.blocks[bi];
  InfBits b;
  b.p = p.comp + blk.in_off; b.end = b.p + blk.in_len; b.buf = 0; b.cnt = 0;
  uint8_t *out = p.out + blk.out_off;
  uint32_t o = 0;
  short lencnt[INF_MAXBITS + 1], lensym[INF_MAXL], distcnt[INF_MAXBITS + 1], distsym[INF_MAXD];
  short lengths[INF_MAXL + INF_MAXD + 2];
  InfHuff lc, dc;
  lc.count = lencnt; lc.symbol = lensym; dc.count = distcnt; dc.symbol = distsym;
  int err = 0, last;
  do {
    last = (int)b.bits(1);
    const int type = (int)b.bits(2);
    if (type == 0) {  // stored
      b.buf >>= (b.cnt & 7); b.cnt &= ~7;  // to the byte boundary
      // the bit buffer holds whole bytes now: give them back
      b.p -= b.cnt >> 3; b.buf = 0; b.cnt = 0;
      if (b.p + 4 > b.end) { err = 6; break; }
      const uint32_t len = (uint32_t)b.p[0] | ((uint32_t)b.p[1] << 8), nlen = (uint32_t)b.p[2] | ((uint32_t)b.p[3] << 8);
      b.p += 4;
      if ((len ^ 0xFFFFu) != nlen || b.p + len > b.end || o + len > blk.out_len) { err = 7; break; }
      for (uint32_t i = 0; i < len; i++) out[o + i] = b.p[i];
      o += len; b.p += len;
    } else if (type == 1) {  // fixed codes
      int s = 0;
      for (; s < 144; s++) lengths[s] = 8;
      for (; s < 256; s++) lengths[s] = 9;
      for (; s < 280; s++) lengths[s] = 7;
      for (; s < 288; s++) lengths[s] = 8;
      inf_construct(lc, lengths, 288);
      for (s = 0; s < 30; s++) lengths[s] = 5;
      inf_construct(dc, lengths, 30);
      err = inf_codes(b, lc, dc, out, o, blk.out_len);
    } else if (type == 2) {  // dynamic codes
      const int nlen = (int)b.bits(5) + 257, ndist = (int)b.bits(5) + 1, ncode = (int)b.bits(4) + 4;
      if (nlen > 286 || ndist > 30) { err = 8; break; }
      int idx = 0;
      for (; idx < ncode; idx++) lengths[INF_ORDER[idx]] = (short)b.bits(3);
      for (; idx < 19; idx++) lengths[INF_ORDER[idx]] = 0;
      if (inf_construct(lc, lengths, 19) != 0) { err = 9; break; }  // the code length code must be complete
      idx = 0;
      while (idx < nlen + ndist) {
        int sym = inf_decode(b, lc);
        if (sym < 0) { err = 10; break; }
        if (sym < 16) {
          lengths[idx++] = (short)sym;
        } else {
          int len = 0, rep;
          if (sym == 16) {
            if (idx == 0) { err = 11; break; }
            len = lengths[idx - 1];
            rep = 3 + (int)b.bits(2);
          } else if (sym == 17) {
            rep = 3 + (int)b.bits(3);
          } else {
            rep = 11 + (int)b.bits(7);
          }
          if (idx + rep > nlen + ndist) { err = 12; break; }
          while (rep--) lengths[idx++] = (short)len;
        }
      }
      if (err) break;
      if (lengths[256] == 0) { err = 13; break; }
      int r = inf_construct(lc, lengths, nlen);
      if (r < 0 || (r > 0 && nlen - lc.count[0] != 1)) { err = 14; break; }
      r = inf_construct(dc, lengths + nlen, ndist);
      if (r < 0 || (r > 0 && ndist - dc.count[0] != 1)) { err = 15; break; }
      err = inf_codes(b, lc, dc, out, o, blk.out_len);
    } else {
      err = 16;
    }
  } while (!last && !err);
  if (err || o != blk.out_len) atomicAdd(p.n_bad, 1u);
}

}  // namespace bvcf
