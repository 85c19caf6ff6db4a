// bvcf_names.cuh -- north-star kernel (4b), sample-name lists of the long rows as aligned 16-byte vectors.
//
// bvcf_tile_kernel (bvcf_tile.cuh) writes the lists of short rows itself and queues the others.  When every
// list item (name + delimiter) has one size of 5..16 bytes and TSV output is on, the queues are served here instead
// of by bvcf_names_big_kernel (7-character names + 1-character delimiter, the 1000 Genomes / biobank layout, are
// 8-byte items with a fast path of their own):
//   sweep   one pass over the row's quad events: het / hom / missing slots as nibble masks (general GT grammar for
//           complex samples, main.go:1126-1190), ranks from one packed warp prefix sum per 64 quads, the sample
//           indices of the three lists compacted into shared memory in header order (main.go:1057 loop order);
//           with --dosageOutput the same pass scatters the int8 dosages (main.go:1172-1178);
//   emit    each list (main.go:617,639,653 strings.Join) leaves as aligned uint4 stores, every vector assembled
//           from the items it overlaps; byte stores only at the ragged ends.
//   bvcf_names_vec_kernel   warp per row (index buffer of 2,816 samples; longer lists in chunks, still one sweep)
//   bvcf_names_long_kernel  CTA per row beyond 4,096 quads: counting sweep, prefix over the warps, chunked sweeps
// The first-generation kernel stored each item with 2-4 narrow unaligned stores (0.6 store sectors per clock per
// SM: the LSU limit, not HBM).
#pragma once
#include "bvcf_rows.cuh"

namespace bvcf {

constexpr int NVEC_WARPS = 2;
constexpr int NVEC_IDX_BYTES = 5632;   // per-warp index buffer: 2816 samples (16-bit) / 1408 (32-bit) per sweep

// `n` names, given by sample index in idx[0, n) (shared memory), as list bytes [g, g + len) of the output,
// len <= 8 n
template <typename IdxT>
__device__ __forceinline__ void emit_name_vectors(const unsigned long long *__restrict__ name8, uint8_t *g, const IdxT *idx,
                                                  uint32_t n, uint32_t len, int lane) {
  const uint32_t head = (16u - (uint32_t)((uintptr_t)g & 15u)) & 15u;
  const uint32_t h = head < len ? head : len;
  for (uint32_t b = lane; b < h; b += 32) g[b] = (uint8_t)(name8[idx[b >> 3]] >> (8 * (b & 7u)));
  if (h == len) return;
  const uint32_t nvec = (len - h) >> 4;
  const uint32_t sh = (h & 7u) * 8u;  // every vector starts at stream offset h + 16 v: same phase within an item
  uint4 *gv = reinterpret_cast<uint4 *>(g + h);
  const uint32_t k00 = h >> 3;
  if (sh == 0) {
    for (uint32_t v = lane; v < nvec; v += 32) {
      const uint32_t k0 = k00 + 2 * v;
      const unsigned long long i0 = name8[idx[k0]];
      const unsigned long long i1 = k0 + 1 < n ? name8[idx[k0 + 1]] : 0ull;
      gv[v] = make_uint4((uint32_t)i0, (uint32_t)(i0 >> 32), (uint32_t)i1, (uint32_t)(i1 >> 32));
    }
  } else {
    for (uint32_t v = lane; v < nvec; v += 32) {
      const uint32_t k0 = k00 + 2 * v;
      const unsigned long long i0 = name8[idx[k0]];
      const unsigned long long i1 = k0 + 1 < n ? name8[idx[k0 + 1]] : 0ull;
      const unsigned long long i2 = k0 + 2 < n ? name8[idx[k0 + 2]] : 0ull;
      const unsigned long long lo = (i0 >> sh) | (i1 << (64u - sh)), hi = (i1 >> sh) | (i2 << (64u - sh));
      gv[v] = make_uint4((uint32_t)lo, (uint32_t)(lo >> 32), (uint32_t)hi, (uint32_t)(hi >> 32));
    }
  }
  for (uint32_t b = h + 16u * nvec + lane; b < len; b += 32) g[b] = (uint8_t)(name8[idx[b >> 3]] >> (8 * (b & 7u)));
}

// ---- any fixed item size -------------------------------------------------------------------------------
// 7-character names with a 1-character delimiter are 8-byte items (name8, emit_name_vectors above).  Any other fixed
// name width with an item of 5..16 bytes uses items zero-padded to 16 bytes: a vector is assembled from the (up to
// four) items it overlaps with 128-bit shifts.
struct ItemTable {
  const unsigned long long *name8;
  const uint4 *name16;
  uint32_t item_bytes;    // I = name width + delimiter length
  uint32_t delim_bytes;   // n names make n I - delim_bytes bytes
};
__device__ __forceinline__ ItemTable item_table(const DevCfg &cfg) {
  ItemTable t;
  t.name8 = cfg.name8; t.name16 = cfg.name16; t.item_bytes = (uint32_t)cfg.item_bytes; t.delim_bytes = (uint32_t)cfg.delim_len;
  return t;
}

template <typename IdxT>
__device__ __forceinline__ void emit_item_vectors(const uint4 *__restrict__ name16, uint32_t I, uint8_t *g, const IdxT *idx,
                                                  uint32_t n, uint32_t len, int lane) {
  auto item_byte = [&](uint32_t b) -> uint8_t {
    const uint32_t k = b / I, o = b - k * I;
    return reinterpret_cast<const uint8_t *>(name16 + idx[k])[o];
  };
  const uint32_t head = (16u - (uint32_t)((uintptr_t)g & 15u)) & 15u;
  const uint32_t h = head < len ? head : len;
  for (uint32_t b = lane; b < h; b += 32) g[b] = item_byte(b);
  if (h == len) return;
  const uint32_t nvec = (len - h) >> 4;
  uint4 *gv = reinterpret_cast<uint4 *>(g + h);
  const uint32_t q512 = 512u / I, r512 = 512u - q512 * I;  // a lane's next vector is 512 bytes on
  uint32_t k = (h + 16u * lane) / I, o = (h + 16u * lane) - k * I;
  for (uint32_t v = lane; v < nvec; v += 32) {
    unsigned __int128 acc = 0;
    uint32_t filled = 0, kk = k, oo = o;
    while (filled < 16u) {
      unsigned __int128 it = 0;
      if (kk < n) {
        const uint4 t = name16[idx[kk]];
        it = ((unsigned __int128)(((unsigned long long)t.w << 32) | t.z) << 64) | (((unsigned long long)t.y << 32) | t.x);
      }
      acc |= (it >> (8u * oo)) << (8u * filled);  // the item from its byte oo on, placed after what is there
      filled += I - oo;
      kk++; oo = 0;
    }
    const unsigned long long lo = (unsigned long long)acc, hi = (unsigned long long)(acc >> 64);
    gv[v] = make_uint4((uint32_t)lo, (uint32_t)(lo >> 32), (uint32_t)hi, (uint32_t)(hi >> 32));
    k += q512; o += r512;
    if (o >= I) { o -= I; k++; }
  }
  for (uint32_t b = h + 16u * nvec + lane; b < len; b += 32) g[b] = item_byte(b);
}

// `n` names by sample index as list bytes [g, g + len), whatever the item size
template <typename IdxT>
__device__ __forceinline__ void emit_items(const ItemTable &t, uint8_t *g, const IdxT *idx, uint32_t n, uint32_t len, int lane) {
  if (t.name8) emit_name_vectors<IdxT>(t.name8, g, idx, n, len, lane);
  else emit_item_vectors<IdxT>(t.name16, t.item_bytes, g, idx, n, len, lane);
}

struct RowEvents {
  const uint32_t *ev;
  uint32_t n_words;
  const uint8_t *L;
  uint32_t content_len, a;
  bool simple;
  int8_t *drow;   // the row of the dosage matrix (zeroed beforehand), or null
};

// int8 dosage of one het / hom / missing sample of a quad: -1 missing, else min(alleles equal to the row's, 127)
// (main.go:1172-1178); `bit` is the sample's nibble flag (bit 4 j + 3)
__device__ __forceinline__ void put_dosage(const RowEvents &re, uint2 e, uint32_t bit, bool is_h, bool is_o, uint32_t samp) {
  int v = -1;
  if (is_h | is_o) {
    if (e.x & EV_COMPLEX) {
      uint32_t gt, alt;
      classify_gt_general(re.L + e.y, re.content_len > e.y ? re.content_len - e.y : 0, re.a, gt, alt);
      v = alt > 127 ? 127 : (int)alt;
    } else {
      const uint32_t sh = (uint32_t)(__ffs(bit) - 1) - 3u;  // 4 * slot
      const bool hap = ((e.y >> (16 + sh)) & 0xFu) == EV_NIB_ABSENT;
      v = is_h ? 1 : (hap ? 1 : 2);
    }
  }
  re.drow[samp] = (int8_t)v;
}

// all three lists of one row in one sweep: index lists at idx[0,n_het) | [n_het, n_het+n_hom) | [.., +n_miss).
// A lane takes two consecutive quads (eight samples) per step, so one packed prefix sum ranks 256 samples.
template <typename IdxT, bool DOSAGE>
__device__ __forceinline__ void index_lists_once(const RowEvents &re, IdxT *idx, uint32_t base_o, uint32_t base_m, int lane) {
  const uint32_t nq = re.n_words >> 1;
  const uint2 *ev2 = reinterpret_cast<const uint2 *>(re.ev);
  uint32_t run_h = 0, run_o = base_o, run_m = base_m;
  uint2 n0 = make_uint2(0u, 0u), n1 = make_uint2(0u, 0u);
  if (2u * lane < nq) n0 = ev2[2 * lane];
  if (2u * lane + 1 < nq) n1 = ev2[2 * lane + 1];
  for (uint32_t base = 0; base < nq; base += 64) {
    const uint2 e0 = n0, e1 = n1;  // software pipelining: the next step's quads are already in flight
    const uint32_t q0 = base + 2 * lane;
    const bool v0 = q0 < nq, v1 = q0 + 1 < nq;
    n0 = make_uint2(0u, 0u); n1 = make_uint2(0u, 0u);
    if (q0 + 64 < nq) n0 = ev2[q0 + 64];
    if (q0 + 65 < nq) n1 = ev2[q0 + 65];
    uint32_t mh0, mo0, mm0, mh1, mo1, mm1;
    quad_masks(e0.x, e0.y, re.a, re.simple, re.L, re.content_len, v0, mh0, mo0, mm0);
    quad_masks(e1.x, e1.y, re.a, re.simple, re.L, re.content_len, v1, mh1, mo1, mm1);
    const uint32_t cnt = (__popc(mh0) + __popc(mh1)) | ((__popc(mo0) + __popc(mo1)) << 10) | ((__popc(mm0) + __popc(mm1)) << 20);
    const uint32_t incl = warp_incl_scan(cnt, lane);
    const uint32_t tot = __shfl_sync(FULL, incl, 31);
    const uint32_t excl = incl - cnt;
    uint32_t kh = run_h + (excl & 1023u), ko = run_o + ((excl >> 10) & 1023u), km = run_m + (excl >> 20);
    const uint32_t s0 = (e0.x & EV_SAMPLE_MASK) - EV_BASE_BIAS, s1 = (e1.x & EV_SAMPLE_MASK) - EV_BASE_BIAS;
    const uint32_t any0 = mh0 | mo0 | mm0, any1 = mh1 | mo1 | mm1;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const uint32_t bit = 8u << (4 * j);
      if (any0 & bit) {
        const uint32_t d = (mh0 & bit) ? kh++ : ((mo0 & bit) ? ko++ : km++);
        idx[d] = (IdxT)(s0 + (uint32_t)j);
        if (DOSAGE && re.drow) put_dosage(re, e0, bit, (mh0 & bit) != 0, (mo0 & bit) != 0, s0 + (uint32_t)j);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const uint32_t bit = 8u << (4 * j);
      if (any1 & bit) {
        const uint32_t d = (mh1 & bit) ? kh++ : ((mo1 & bit) ? ko++ : km++);
        idx[d] = (IdxT)(s1 + (uint32_t)j);
        if (DOSAGE && re.drow) put_dosage(re, e1, bit, (mh1 & bit) != 0, (mo1 & bit) != 0, s1 + (uint32_t)j);
      }
    }
    run_h += tot & 1023u; run_o += (tot >> 10) & 1023u; run_m += tot >> 20;
  }
}

// Rows with more names than the index buffer holds (wide cohorts): still ONE sweep over the quads.  The buffer
// is cut into three regions; a list whose region cannot take another step (256 names) is written out as a chunk
// of vectors and its region starts over.  n names make 8 n - 1 bytes: the final chunk of a list drops the
// trailing delimiter.
template <typename IdxT, bool DOSAGE>
__device__ __forceinline__ void sweep_lists_chunked(const ItemTable &tab, const RowEvents &re,
                                                    IdxT *idx, uint32_t cap3, uint8_t *g_h, uint8_t *g_o, uint8_t *g_m,
                                                    uint32_t n_h, uint32_t n_o, uint32_t n_m, int lane,
                                                    uint32_t sh = 0, uint32_t so = 0, uint32_t sm = 0) {
  // sh, so, sm: names of each list before this stretch of quads (a CTA splits a very long row among its warps)
  const uint32_t nq = re.n_words >> 1;
  const uint2 *ev2 = reinterpret_cast<const uint2 *>(re.ev);
  IdxT *const ih = idx, *const io = idx + cap3, *const im = idx + 2 * cap3;
  uint32_t fh = 0, fo = 0, fm = 0;          // names waiting in each region
  uint2 n0 = make_uint2(0u, 0u), n1 = make_uint2(0u, 0u);
  if (2u * lane < nq) n0 = ev2[2 * lane];
  if (2u * lane + 1 < nq) n1 = ev2[2 * lane + 1];
  for (uint32_t base = 0; base < nq; base += 64) {
    const uint2 e0 = n0, e1 = n1;
    const uint32_t q0 = base + 2 * lane;
    const bool v0 = q0 < nq, v1 = q0 + 1 < nq;
    n0 = make_uint2(0u, 0u); n1 = make_uint2(0u, 0u);
    if (q0 + 64 < nq) n0 = ev2[q0 + 64];
    if (q0 + 65 < nq) n1 = ev2[q0 + 65];
    uint32_t mh0, mo0, mm0, mh1, mo1, mm1;
    quad_masks(e0.x, e0.y, re.a, re.simple, re.L, re.content_len, v0, mh0, mo0, mm0);
    quad_masks(e1.x, e1.y, re.a, re.simple, re.L, re.content_len, v1, mh1, mo1, mm1);
    const uint32_t cnt = (__popc(mh0) + __popc(mh1)) | ((__popc(mo0) + __popc(mo1)) << 10) | ((__popc(mm0) + __popc(mm1)) << 20);
    const uint32_t incl = warp_incl_scan(cnt, lane);
    const uint32_t tot = __shfl_sync(FULL, incl, 31);
    const uint32_t excl = incl - cnt;
    uint32_t kh = fh + (excl & 1023u), ko = fo + ((excl >> 10) & 1023u), km = fm + (excl >> 20);
    const uint32_t s0 = (e0.x & EV_SAMPLE_MASK) - EV_BASE_BIAS, s1 = (e1.x & EV_SAMPLE_MASK) - EV_BASE_BIAS;
    const uint32_t any0 = mh0 | mo0 | mm0, any1 = mh1 | mo1 | mm1;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const uint32_t bit = 8u << (4 * j);
      if (any0 & bit) {
        if (mh0 & bit) ih[kh++] = (IdxT)(s0 + (uint32_t)j);
        else if (mo0 & bit) io[ko++] = (IdxT)(s0 + (uint32_t)j);
        else im[km++] = (IdxT)(s0 + (uint32_t)j);
        if (DOSAGE && re.drow) put_dosage(re, e0, bit, (mh0 & bit) != 0, (mo0 & bit) != 0, s0 + (uint32_t)j);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const uint32_t bit = 8u << (4 * j);
      if (any1 & bit) {
        if (mh1 & bit) ih[kh++] = (IdxT)(s1 + (uint32_t)j);
        else if (mo1 & bit) io[ko++] = (IdxT)(s1 + (uint32_t)j);
        else im[km++] = (IdxT)(s1 + (uint32_t)j);
        if (DOSAGE && re.drow) put_dosage(re, e1, bit, (mh1 & bit) != 0, (mo1 & bit) != 0, s1 + (uint32_t)j);
      }
    }
    const uint32_t th = tot & 1023u, to = (tot >> 10) & 1023u, tm = tot >> 20;
    fh += th; fo += to; fm += tm;
    sh += th; so += to; sm += tm;
    const bool last = base + 64 >= nq;
    if (last || fh + 256 > cap3 || fo + 256 > cap3 || fm + 256 > cap3) {
      __syncwarp();
      if (fh && (last || fh + 256 > cap3)) {
        const uint32_t len = tab.item_bytes * fh - (sh >= n_h ? tab.delim_bytes : 0u);
        emit_items<IdxT>(tab, g_h, ih, fh, len, lane);
        g_h += len; fh = 0;
      }
      if (fo && (last || fo + 256 > cap3)) {
        const uint32_t len = tab.item_bytes * fo - (so >= n_o ? tab.delim_bytes : 0u);
        emit_items<IdxT>(tab, g_o, io, fo, len, lane);
        g_o += len; fo = 0;
      }
      if (fm && (last || fm + 256 > cap3)) {
        const uint32_t len = tab.item_bytes * fm - (sm >= n_m ? tab.delim_bytes : 0u);
        emit_items<IdxT>(tab, g_m, im, fm, len, lane);
        g_m += len; fm = 0;
      }
      __syncwarp();
    }
  }
}

template <typename IdxT, bool DOSAGE>
__device__ __forceinline__ void names_row_vec(const NamesParams &p, const RowDesc &rd, IdxT *idx, int lane) {
  const DevCfg &cfg = p.cfg;
  const LineRec rec = p.lines[rd.line];
  RowEvents re;
  re.ev = p.events + rec.ev_start; re.n_words = rec.ev_count; re.L = p.in + rec.start;
  re.content_len = rec.len >= (uint32_t)cfg.eol_width ? rec.len - (uint32_t)cfg.eol_width : 0;
  re.a = rd.allele; re.simple = !(rec.flags & 1) && rd.allele == 1;
  re.drow = nullptr;
  {
    const unsigned long long gr = p.ctr->chunk_row_base + rd.row;
    if (DOSAGE && cfg.want_dosage && gr < p.dosage_cap_rows) re.drow = p.dosage + gr * (unsigned long long)cfg.n_samples;
  }
  const ItemTable tab = item_table(cfg);
  const uint32_t cap = NVEC_IDX_BYTES / sizeof(IdxT);
  const uint32_t n_tot = rd.n_het + rd.n_hom + rd.n_miss;
  const uint32_t ns[3] = {rd.n_het, rd.n_hom, rd.n_miss};
  const unsigned long long dsts[3] = {rd.het_dst, rd.hom_dst, rd.miss_dst};
  if (n_tot <= cap) {
    index_lists_once<IdxT, DOSAGE>(re, idx, rd.n_het, rd.n_het + rd.n_hom, lane);
    __syncwarp();
    uint32_t b = 0;
#pragma unroll
    for (int c = 0; c < 3; c++) {
      if (ns[c]) emit_items<IdxT>(tab, p.out + dsts[c], idx + b, ns[c], tab.item_bytes * ns[c] - tab.delim_bytes, lane);
      b += ns[c];
    }
    __syncwarp();
  } else {
    sweep_lists_chunked<IdxT, DOSAGE>(tab, re, idx, cap / 3, p.out + rd.het_dst, p.out + rd.hom_dst, p.out + rd.miss_dst, rd.n_het,
                              rd.n_hom, rd.n_miss, lane);
  }
}

// warp per queued row (requires cfg.name8 and TSV output; also scatters the int8 dosage row when one is wanted)
template <typename IdxT, bool DOSAGE>
__global__ void __launch_bounds__(NVEC_WARPS * 32) bvcf_names_vec_kernel(const NamesParams p) {
  __shared__ __align__(16) uint8_t s_idx[NVEC_WARPS][NVEC_IDX_BYTES];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (p.ctr->out_overflow | p.ctr->ev_overflow | p.ctr->slot_overflow | p.ctr->row_overflow) return;
  const uint32_t n_mid = p.ctr->n_mid_rows, n_big = p.ctr->n_big_rows;  // this kernel's rows follow the mid list
  IdxT *idx = reinterpret_cast<IdxT *>(s_idx[warp]);
  // rows differ by three orders of magnitude in size: warps take the next queued row from a shared cursor
  for (;;) {
    uint32_t wi = 0;
    if (lane == 0) wi = atomicAdd(&p.ctr->big_row_cursor, 1u);
    wi = __shfl_sync(FULL, wi, 0);
    if (wi >= n_big) break;
    names_row_vec<IdxT, DOSAGE>(p, p.row_desc[n_mid + wi], idx, lane);
  }
}

// CTA per very long row (biobank width: tens of thousands of quads, up to megabytes of names).  Each warp takes a
// contiguous stretch of the row's quads: a counting sweep, a prefix over the eight warps, then the chunked sweep
// writes the warp's part of the three lists at its offsets.
constexpr int NLONG_WARPS = 8;
constexpr int NLONG_IDX_BYTES = 5120;
template <typename IdxT, bool DOSAGE>
__global__ void __launch_bounds__(NLONG_WARPS * 32) bvcf_names_long_kernel(const NamesParams p) {
  __shared__ __align__(16) uint8_t s_idx[NLONG_WARPS][NLONG_IDX_BYTES];
  __shared__ uint32_t s_cnt[NLONG_WARPS][3];
  __shared__ uint32_t s_wi;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (p.ctr->out_overflow | p.ctr->ev_overflow | p.ctr->slot_overflow | p.ctr->row_overflow) return;
  const DevCfg &cfg = p.cfg;
  const uint32_t n_long = p.ctr->n_long_rows;
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) s_wi = atomicAdd(&p.ctr->long_row_cursor, 1u);
    __syncthreads();
    const uint32_t wi = s_wi;
    if (wi >= n_long) break;
    const RowDesc rd = p.row_desc[p.row_desc_cap - 1 - wi];
    const LineRec rec = p.lines[rd.line];
    const uint32_t nq = rec.ev_count >> 1;
    const uint32_t seg = ((nq + NLONG_WARPS - 1) / NLONG_WARPS + 63u) & ~63u;  // whole 64-quad steps
    const uint32_t q_lo = warp * seg < nq ? warp * seg : nq, q_hi = q_lo + seg < nq ? q_lo + seg : nq;
    RowEvents re;
    re.ev = p.events + rec.ev_start + 2 * q_lo; re.n_words = 2 * (q_hi - q_lo); re.L = p.in + rec.start;
    re.content_len = rec.len >= (uint32_t)cfg.eol_width ? rec.len - (uint32_t)cfg.eol_width : 0;
    re.a = rd.allele; re.simple = !(rec.flags & 1) && rd.allele == 1;
    re.drow = nullptr;
    {
      const unsigned long long gr = p.ctr->chunk_row_base + rd.row;
      if (DOSAGE && cfg.want_dosage && gr < p.dosage_cap_rows) re.drow = p.dosage + gr * (unsigned long long)cfg.n_samples;
    }
    // counting sweep
    uint32_t ch = 0, co = 0, cm = 0;
    const uint2 *ev2 = reinterpret_cast<const uint2 *>(re.ev);
    for (uint32_t q = lane; q < q_hi - q_lo; q += 32) {
      const uint2 e = ev2[q];
      uint32_t mh, mo, mm;
      quad_masks(e.x, e.y, re.a, re.simple, re.L, re.content_len, true, mh, mo, mm);
      ch += __popc(mh); co += __popc(mo); cm += __popc(mm);
    }
    ch = __reduce_add_sync(FULL, ch); co = __reduce_add_sync(FULL, co); cm = __reduce_add_sync(FULL, cm);
    if (lane == 0) { s_cnt[warp][0] = ch; s_cnt[warp][1] = co; s_cnt[warp][2] = cm; }
    __syncthreads();
    uint32_t ph = 0, po = 0, pm = 0;
    for (int w = 0; w < warp; w++) { ph += s_cnt[w][0]; po += s_cnt[w][1]; pm += s_cnt[w][2]; }
    const uint32_t cap3 = (NLONG_IDX_BYTES / sizeof(IdxT)) / 3;
    const ItemTable tab = item_table(cfg);
    const unsigned long long I = tab.item_bytes;
    sweep_lists_chunked<IdxT, DOSAGE>(tab, re, reinterpret_cast<IdxT *>(s_idx[warp]), cap3, p.out + rd.het_dst + I * ph,
                              p.out + rd.hom_dst + I * po, p.out + rd.miss_dst + I * pm, rd.n_het, rd.n_hom, rd.n_miss,
                              lane, ph, po, pm);
  }
}

}  // namespace bvcf
