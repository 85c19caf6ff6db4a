// bvcf_synth.cu -- libbvcfsynth.so: generates the synthetic benchmark workloads on the host or straight into
// device memory.  Bench/test infrastructure; the transform library (libbvcf.so) does not depend on it.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "bvcf_synth.cuh"

using namespace bvcf_synth;

namespace {

constexpr int PREFIX_CAP = 768;

__global__ void synth_len_kernel(Params P, uint64_t first, uint64_t n, uint32_t *lens) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint8_t buf[PREFIX_CAP];
  LineGeno g;
  const uint32_t pl = line_prefix(P, first + i, buf, g);
  lens[i] = (uint32_t)line_len(P, pl);
}

// exclusive scan of lens -> offs (single block per 1M elements is plenty for a one-off generator)
__global__ void synth_scan_kernel(const uint32_t *lens, uint64_t *offs, uint64_t n, uint64_t *total) {
  __shared__ unsigned long long sh[1024];
  __shared__ unsigned long long carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (uint64_t base = 0; base < n; base += 1024 * 8) {
    unsigned long long v[8], s = 0;
    for (int k = 0; k < 8; k++) {
      const uint64_t i = base + (uint64_t)threadIdx.x * 8 + k;
      v[k] = i < n ? lens[i] : 0;
      s += v[k];
    }
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {
      unsigned long long x = threadIdx.x >= d ? sh[threadIdx.x - d] : 0;
      __syncthreads();
      sh[threadIdx.x] += x;
      __syncthreads();
    }
    unsigned long long run = carry + sh[threadIdx.x] - s;
    for (int k = 0; k < 8; k++) {
      const uint64_t i = base + (uint64_t)threadIdx.x * 8 + k;
      if (i < n) offs[i] = run;
      run += v[k];
    }
    __syncthreads();
    if (threadIdx.x == 1023) carry += sh[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry;
}

// one warp per line: lane 0 renders the fixed fields into shared memory, all lanes write bytes
__global__ void __launch_bounds__(256) synth_fill_kernel(Params P, uint64_t first, uint64_t n, const uint64_t *offs,
                                                          uint8_t *out) {
  __shared__ uint8_t spre[8][PREFIX_CAP];
  __shared__ LineGeno sg[8];
  __shared__ uint32_t spl[8];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint64_t nw = (uint64_t)gridDim.x * 8;
  for (uint64_t i = (uint64_t)blockIdx.x * 8 + w; i < n; i += nw) {
    __syncwarp();
    if (lane == 0) {
      LineGeno g;
      spl[w] = line_prefix(P, first + i, spre[w], g);
      sg[w] = g;
    }
    __syncwarp();
    const uint32_t pl = spl[w];
    const LineGeno g = sg[w];
    uint8_t *dst = out + offs[i];
    for (uint32_t k = lane; k < pl; k += 32) dst[k] = spre[w][k];
    dst += pl;
    const uint64_t nb = 4ull * P.n_samples;
    for (uint64_t j = lane; j < nb; j += 32) dst[j] = gt_byte(P, g, first + i, j);
  }
}

void host_fill(const Params &P, uint64_t first, uint64_t n, const uint64_t *offs, uint8_t *out) {
  uint8_t pre[PREFIX_CAP];
  for (uint64_t i = 0; i < n; i++) {
    LineGeno g;
    const uint32_t pl = line_prefix(P, first + i, pre, g);
    uint8_t *dst = out + offs[i];
    memcpy(dst, pre, pl);
    dst += pl;
    const uint64_t nb = 4ull * P.n_samples;
    for (uint64_t j = 0; j < nb; j++) dst[j] = gt_byte(P, g, first + i, j);
  }
}

}  // namespace

extern "C" {

// "##fileformat" + meta lines + "#CHROM ..." header; sample names [A-Z]{2}[0-9]{5} (no dots).
// Returns the length; writes at most cap bytes.
size_t bvcf_synth_header(uint64_t seed, uint32_t n_samples, char *buf, size_t cap) {
  std::string h = "##fileformat=VCFv4.1\n##source=bvcf_synth\n##INFO=<ID=AC,Number=A,Type=Integer,Description=\"alt count\">\n"
                  "##FORMAT=<ID=GT,Number=1,Type=String,Description=\"Genotype\">\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO";
  if (n_samples) h += "\tFORMAT";
  for (uint32_t s = 0; s < n_samples; s++) {
    const uint64_t x = mix64(seed * 31 + 7 + s);
    char nm[16];
    snprintf(nm, sizeof nm, "\t%c%c%05u", 'A' + (int)(x % 26), 'A' + (int)((x >> 8) % 26), (unsigned)(s % 100000));
    h += nm;
  }
  h += "\n";
  if (buf && cap) memcpy(buf, h.data(), std::min(cap, h.size()));
  return h.size();
}

// Bytes that lines [first, first+n) occupy (host computation, multi-threaded).
uint64_t bvcf_synth_host_size(uint64_t seed, uint32_t n_samples, int shape, uint64_t first, uint64_t n) {
  Params P{seed, n_samples, shape};
  uint64_t total = 0;
  uint8_t pre[PREFIX_CAP];
  for (uint64_t i = 0; i < n; i++) {
    LineGeno g;
    total += line_len(P, line_prefix(P, first + i, pre, g));
  }
  return total;
}

// Generate lines [first, first+n) into `out` on the host.  Returns bytes written, or 0 if cap is too small.
uint64_t bvcf_synth_host(uint64_t seed, uint32_t n_samples, int shape, uint64_t first, uint64_t n, uint8_t *out,
                         uint64_t cap, int threads) {
  Params P{seed, n_samples, shape};
  std::vector<uint64_t> offs(n + 1);
  uint8_t pre[PREFIX_CAP];
  uint64_t run = 0;
  for (uint64_t i = 0; i < n; i++) {
    LineGeno g;
    offs[i] = run;
    run += line_len(P, line_prefix(P, first + i, pre, g));
  }
  offs[n] = run;
  if (run > cap) return 0;
  if (threads < 1) threads = 1;
  std::vector<std::thread> th;
  for (int t = 0; t < threads; t++) {
    const uint64_t lo = n * t / threads, hi = n * (t + 1) / threads;
    th.emplace_back([&, lo, hi] { host_fill(P, first + lo, hi - lo, offs.data() + lo, out); });
  }
  for (auto &t : th) t.join();
  return run;
}

// Generate lines [first, first+n) straight into device memory d_out (capacity cap bytes) on `device`.
// Returns bytes written; 0 on error / insufficient capacity (then *needed holds the size).
uint64_t bvcf_synth_device(uint64_t seed, uint32_t n_samples, int shape, uint64_t first, uint64_t n, void *d_out,
                           uint64_t cap, int device, uint64_t *needed) {
  Params P{seed, n_samples, shape};
  if (cudaSetDevice(device) != cudaSuccess) return 0;
  uint32_t *lens = nullptr;
  uint64_t *offs = nullptr, *d_total = nullptr;
  uint64_t total = 0;
  if (cudaMalloc(&lens, n * 4 + 16) != cudaSuccess) return 0;
  if (cudaMalloc(&offs, n * 8 + 16) != cudaSuccess) { cudaFree(lens); return 0; }
  if (cudaMalloc(&d_total, 8) != cudaSuccess) { cudaFree(lens); cudaFree(offs); return 0; }
  synth_len_kernel<<<(unsigned)((n + 127) / 128), 128>>>(P, first, n, lens);
  synth_scan_kernel<<<1, 1024>>>(lens, offs, n, d_total);
  cudaMemcpy(&total, d_total, 8, cudaMemcpyDeviceToHost);
  if (needed) *needed = total;
  uint64_t ret = 0;
  if (total <= cap) {
    synth_fill_kernel<<<148 * 8, 256>>>(P, first, n, offs, (uint8_t *)d_out);
    if (cudaDeviceSynchronize() == cudaSuccess && cudaGetLastError() == cudaSuccess) ret = total;
  }
  cudaFree(lens);
  cudaFree(offs);
  cudaFree(d_total);
  return ret;
}

}  // extern "C"
