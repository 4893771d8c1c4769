// ipb200 — the host side of the host-pointer pipeline (no CUDA in here): rebuilding the caller's LOGICAL*1 array from the
// packed lo the device sends (ipb_core.cuh: k_pack_lo), and row copies between the caller's pageable arrays and the
// library's pinned staging, both on a few host threads.  ipgpu_selftest_host (ipb_api.cu) exposes them to the CPU tests.
#pragma once
#include <algorithm>
#include <cstring>
#include <thread>
#include <vector>

#include "ipb_grid.h"

namespace ipb {

inline void expand_lo_fields(u8* lo, i64 mo, i64 no, i64 kbeg, i64 kend, i64 kstep, const unsigned* bits, i64 nw, const int* hasfalse) {
  static const std::vector<unsigned long long> lut = [] {
    std::vector<unsigned long long> t(256);
    for (int b = 0; b < 256; b++) { unsigned long long v = 0; for (int j = 0; j < 8; j++) v |= (unsigned long long)((b >> j) & 1) << (8 * j); t[b] = v; }
    return t;
  }();
  for (i64 k = kbeg; k < kend; k += kstep) {
    u8* o = lo + (size_t)k * mo;
    if (!hasfalse[k]) { memset(o, 1, (size_t)no); continue; }
    const unsigned char* b = reinterpret_cast<const unsigned char*>(bits + (size_t)k * nw);
    const i64 full = no / 8;
    for (i64 q = 0; q < full; q++) memcpy(o + 8 * q, &lut[b[q]], 8);   // little-endian hosts (x86-64, aarch64)
    for (i64 n = full * 8; n < no; n++) o[n] = (u8)((b[n >> 3] >> (n & 7)) & 1);
  }
}
inline void expand_lo(u8* lo, i64 mo, i64 no, i64 kn, const unsigned* bits, i64 nw, const int* hasfalse, int threads) {
  const int T = (int)std::max<i64>(1, std::min<i64>(threads, (kn * no) >> 20));   // a thread per MiB of lo at least
  if (T <= 1) { expand_lo_fields(lo, mo, no, 0, kn, 1, bits, nw, hasfalse); return; }
  std::vector<std::thread> th;
  for (int t = 1; t < T; t++) th.emplace_back(expand_lo_fields, lo, mo, no, (i64)t, kn, (i64)T, bits, nw, hasfalse);
  expand_lo_fields(lo, mo, no, 0, kn, T, bits, nw, hasfalse);
  for (auto& x : th) x.join();
}

// rows of `width` bytes from src (pitch spitch) to dst (pitch dpitch) on `threads` host threads: the byte range is cut
// into equal pieces, whatever the row count
inline void copy_rows_range(unsigned char* dst, size_t dpitch, const unsigned char* src, size_t spitch, size_t width, size_t b0, size_t b1) {
  while (b0 < b1) {
    const size_t r = b0 / width, off = b0 - r * width, n = std::min(width - off, b1 - b0);
    memcpy(dst + r * dpitch + off, src + r * spitch + off, n);
    b0 += n;
  }
}
inline void copy_rows(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t rows, int threads) {
  const size_t total = width * rows;
  if (total == 0) return;
  const int T = (int)std::max<size_t>(1, std::min<size_t>((size_t)threads, total >> 20));
  unsigned char* d = static_cast<unsigned char*>(dst);
  const unsigned char* q = static_cast<const unsigned char*>(src);
  if (T <= 1) { copy_rows_range(d, dpitch, q, spitch, width, 0, total); return; }
  const size_t piece = (total + T - 1) / T;
  std::vector<std::thread> th;
  for (int t = 1; t < T; t++) th.emplace_back(copy_rows_range, d, dpitch, q, spitch, width, std::min(total, t * piece), std::min(total, (t + 1) * piece));
  copy_rows_range(d, dpitch, q, spitch, width, 0, std::min(total, piece));
  for (auto& x : th) x.join();
}

}  // namespace ipb
