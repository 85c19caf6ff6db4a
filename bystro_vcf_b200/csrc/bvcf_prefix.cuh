// bvcf_prefix.cuh -- device-wide exclusive prefix sums (hand-written, fixed grid, device-side n) and the
// line-table compaction that puts the scan kernel's per-range records into input order.
//
// The element count lives in device memory (RunCounters::chunk_records) so the host never has to
// synchronise between the pipeline stages of a sub-chunk.  Three launches: per-span reduce, one-block scan of
// the span totals (also advances the run's output cursors), per-span scan.
#pragma once
#include "bvcf_common.cuh"

namespace bvcf {

constexpr int PFX_BLOCKS = 592;   // 4 x 148 SMs
constexpr int PFX_THREADS = 256;

struct PrefixParams {
  const uint32_t *a, *b;        // two input arrays scanned together (b may be null)
  uint64_t *out_a, *out_b;      // exclusive prefix sums
  unsigned long long *partial;  // 2 * PFX_BLOCKS
  const unsigned int *n_ptr;    // element count in device memory, or null
  uint32_t n_imm;               // used when n_ptr == null
  RunCounters *ctr;
};

__device__ __forceinline__ uint32_t pfx_n(const PrefixParams &p) { return p.n_ptr ? *p.n_ptr : p.n_imm; }

__device__ __forceinline__ void pfx_span(uint32_t n, uint32_t &lo, uint32_t &hi) {
  const uint32_t span = (n + PFX_BLOCKS - 1) / PFX_BLOCKS;
  const unsigned long long l = (unsigned long long)blockIdx.x * span;
  lo = l > n ? n : (uint32_t)l;
  hi = l + span > n ? n : (uint32_t)(l + span);
}

__device__ __forceinline__ unsigned long long block_sum64(unsigned long long v, unsigned long long *sh) {
  v = warp_sum64(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  unsigned long long t = 0;
  for (int i = 0; i < (int)(blockDim.x >> 5); i++) t += sh[i];
  return t;
}

__global__ void __launch_bounds__(PFX_THREADS) bvcf_prefix_reduce_kernel(const PrefixParams p) {
  __shared__ unsigned long long sh[32];
  uint32_t lo, hi;
  pfx_span(pfx_n(p), lo, hi);
  unsigned long long sa = 0, sb = 0;
  for (uint32_t i = lo + threadIdx.x; i < hi; i += PFX_THREADS) {
    sa += p.a[i];
    if (p.b) sb += p.b[i];
  }
  sa = block_sum64(sa, sh);
  sb = block_sum64(sb, sh);
  if (threadIdx.x == 0) {
    p.partial[2 * blockIdx.x] = sa;
    p.partial[2 * blockIdx.x + 1] = sb;
  }
}

// one block: exclusive scan of the PFX_BLOCKS span totals; publishes the totals
__global__ void __launch_bounds__(1024) bvcf_prefix_spine_kernel(const PrefixParams p) {
  __shared__ unsigned long long sa[1024], sb[1024];
  const int t = threadIdx.x;
  unsigned long long va = t < PFX_BLOCKS ? p.partial[2 * t] : 0, vb = t < PFX_BLOCKS ? p.partial[2 * t + 1] : 0;
  sa[t] = va; sb[t] = vb;
  __syncthreads();
  for (int d = 1; d < 1024; d <<= 1) {
    unsigned long long xa = t >= d ? sa[t - d] : 0, xb = t >= d ? sb[t - d] : 0;
    __syncthreads();
    sa[t] += xa; sb[t] += xb;
    __syncthreads();
  }
  if (t < PFX_BLOCKS) {
    p.partial[2 * t] = sa[t] - va;
    p.partial[2 * t + 1] = sb[t] - vb;
  }
  if (t == 1023) {
    RunCounters *c = p.ctr;
    // per-range counts of one sub-chunk: publish the record count, start the sub-chunk's cursors and work lists
    c->chunk_records = (unsigned int)sa[t];
    c->chunk_line_base = c->n_lines;
    c->n_records += sa[t];
    c->n_lines += sb[t];
    c->n_big_recs = 0; c->big_rec_cursor = 0;
    c->tile_ticket = 0; c->tile_ticket2 = 0; c->n_slow = 0; c->scratch_cursor = 0;
    c->n_mid_rows = 0; c->n_big_rows = 0; c->big_row_cursor = 0;
    c->n_long_rows = 0; c->long_row_cursor = 0;
    c->chunk_out_base = c->out_cursor;
    c->chunk_row_base = c->row_cursor;
    c->chunk_loci_base = c->loci_cursor;
  }
}

__global__ void __launch_bounds__(PFX_THREADS) bvcf_prefix_scan_kernel(const PrefixParams p) {
  __shared__ unsigned long long wa[PFX_THREADS / 32], wb[PFX_THREADS / 32];
  uint32_t lo, hi;
  pfx_span(pfx_n(p), lo, hi);
  unsigned long long run_a = p.partial[2 * blockIdx.x], run_b = p.partial[2 * blockIdx.x + 1];
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  for (uint32_t base = lo; base < hi; base += PFX_THREADS) {
    const uint32_t i = base + threadIdx.x;
    const unsigned long long va = i < hi ? p.a[i] : 0, vb = (i < hi && p.b) ? p.b[i] : 0;
    unsigned long long ia = va, ib = vb;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned long long ta = __shfl_up_sync(FULL, ia, d), tb = __shfl_up_sync(FULL, ib, d);
      if (l >= d) { ia += ta; ib += tb; }
    }
    __syncthreads();
    if (l == 31) { wa[w] = ia; wb[w] = ib; }
    __syncthreads();
    unsigned long long oa = 0, ob = 0, ta = 0, tb = 0;
#pragma unroll
    for (int k = 0; k < PFX_THREADS / 32; k++) {
      if (k < w) { oa += wa[k]; ob += wb[k]; }
      ta += wa[k]; tb += wb[k];
    }
    if (i < hi) {
      p.out_a[i] = run_a + oa + ia - va;
      if (p.out_b) p.out_b[i] = run_b + ob + ib - vb;
    }
    run_a += ta; run_b += tb;
  }
}

// ---- compaction: per-range record slots -> dense, input-ordered line table ---------------------
struct CompactParams {
  const LineRec *recs;          // n_ranges * slots_per_range
  const uint32_t *range_nrec;
  const uint64_t *rec_base;     // exclusive prefix of range_nrec
  const uint64_t *line_base;    // exclusive prefix of range_nlines
  LineRec *dense;
  uint32_t n_ranges, slots_per_range, evcap_words;
};

__global__ void __launch_bounds__(256) bvcf_compact_lines_kernel(const CompactParams p) {
  const uint32_t r = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= p.n_ranges) return;
  uint32_t n = p.range_nrec[r];
  if (n > p.slots_per_range) n = p.slots_per_range;  // overflow is flagged; stay in bounds
  const LineRec *src = p.recs + (size_t)r * p.slots_per_range;
  LineRec *dst = p.dense + p.rec_base[r];
  const uint32_t lb = (uint32_t)p.line_base[r];
  for (uint32_t j = lane; j < n; j += 32) {
    LineRec x = src[j];
    x.ev_start += r * p.evcap_words;  // slice-relative -> sub-chunk-relative
    x.ord += lb;
    dst[j] = x;
  }
}

}  // namespace bvcf
