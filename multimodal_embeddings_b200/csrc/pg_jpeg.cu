// pg_jpeg.cu — D1-D8: baseline JPEG scans decoded on the device ("image decode -> HBM").
//
// Replaces the `cv2.imread(image_path)` that opens the path (1_doclayout_bboxes.py:381; again per grid, and at
// 2_edge_box_filter.py:195) for `.jpg` input: the file bytes cross PCIe compressed (a broadsheet scan is
// 5-10 MB as JPEG against 144 MB as the BGR array cv2 returns) and the page is produced in HBM, where the
// tiler reads it.  Greyscale files decode to ONE plane (cv2 would replicate it three times); the tiler's
// one-channel variant reads that plane and writes the same three output planes.
//
// Host side (this file, C++): marker parsing, Huffman/quantisation tables.  Device side: entropy-coded
// segment unstuffed (D1-D3), self-synchronising chunked Huffman decode (D4 speculative pass, D5 sync rounds,
// D6 segmented scan, D7 coefficient store; algorithm in pg_jpeg.h), inverse DCT (D8).  No tensor cores: the
// IDCT is 8x8 integer butterflies with libjpeg's 13-bit constants and must be bit-exact.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "pg_common.cuh"
#include "pg_jpeg.h"

namespace {

constexpr int UB_BYTES = 4096;      // unstuff: input bytes per CTA (256 threads x 16)
constexpr int CHUNK_THREADS = 256;  // entropy kernels: chunks per CTA
constexpr int STREAM_PAD = PGJ_STREAM_PAD;  // readable slack behind every unstuffed stream
constexpr int MAX_ROUNDS = 64;

static_assert(sizeof(PgjImage) % 16 == 0, "PgjImage is copied into shared memory 16 bytes at a time");

struct HostImage {
  PgjImage dev;
  int64_t scan_begin = 0, scan_end = 0;  // byte offsets inside the file
  int32_t n_intervals = 1;
};

// ---- marker parsing (host) ------------------------------------------------------------------------------
// EXIF orientation (APP1 "Exif", TIFF tag 0x0112): cv2.imread turns the decoded image accordingly; a file that asks
// for anything but the identity is left to the host decoder.  Returns the tag's value, 1 when absent / unreadable.
int exif_orientation(const uint8_t* s, int n) {
  if (n < 14 || std::memcmp(s, "Exif\0\0", 6) != 0) return 1;
  const uint8_t* t = s + 6;
  const int m = n - 6;
  const bool le = t[0] == 'I' && t[1] == 'I', be = t[0] == 'M' && t[1] == 'M';
  if (!le && !be) return 1;
  auto u16 = [&](int o) { return le ? (t[o] | t[o + 1] << 8) : (t[o] << 8 | t[o + 1]); };
  auto u32 = [&](int o) { return le ? (uint32_t)(t[o] | t[o + 1] << 8 | t[o + 2] << 16) | (uint32_t)t[o + 3] << 24
                                    : (uint32_t)(t[o + 3] | t[o + 2] << 8 | t[o + 1] << 16) | (uint32_t)t[o] << 24; };
  if (u16(2) != 42) return 1;
  const uint32_t ifd = u32(4);
  if (ifd + 2 > (uint32_t)m) return 1;
  const int cnt = u16((int)ifd);
  for (int i = 0; i < cnt; ++i) {
    const uint32_t e = ifd + 2 + 12u * (uint32_t)i;
    if (e + 12 > (uint32_t)m) return 1;
    if (u16((int)e) == 0x0112) return u16((int)e + 8);
  }
  return 1;
}

int build_huff(const uint8_t* counts, const uint8_t* symbols, int n_symbols, PgjHuff& h) {
  std::memset(&h, 0, sizeof(h));
  int code = 0, k = 0;
  for (int l = 1; l <= 16; ++l) {
    h.valoff[l] = k - code;
    for (int i = 0; i < counts[l - 1]; ++i) {
      if (k >= n_symbols || k >= 256) return -1;
      if (code >= (1 << l)) return -1;  // more codes of this length than the code space holds (the views below index by code)
      if (l <= PGJ_LUT_BITS) {
        const int first = code << (PGJ_LUT_BITS - l), n = 1 << (PGJ_LUT_BITS - l);
        const int sym = symbols[k], run = sym >> 4, size = sym & 15;
        for (int j = 0; j < n; ++j) {
          h.lut[first + j] = (uint16_t)((l << 8) | sym);
          // AC views of the same slot (unused for DC tables)
          h.skip[first + j] = (uint16_t)((l + size) | (run << 5) | (size == 0 ? 512 : 0));
          if (size != 0 && l + size <= PGJ_LUT_BITS) {
            // the magnitude bits follow the code inside the 9-bit window
            const int rest = (first + j) & ((1 << (PGJ_LUT_BITS - l)) - 1);
            const int mag = rest >> (PGJ_LUT_BITS - l - size);
            const int val = mag < (1 << (size - 1)) ? mag - (1 << size) + 1 : mag;
            if (val >= -128 && val <= 127) h.fast[first + j] = (int16_t)(val * 256 + run * 16 + l + size);
          }
        }
      }
      h.vals[k] = symbols[k];
      ++code;
      ++k;
    }
    h.maxcode[l] = counts[l - 1] ? code - 1 : -1;
    if (code > (1 << l)) return -1;
    code <<= 1;
  }
  h.maxcode[17] = 0x7fffffff;
  return 0;
}

int parse_jpeg(const uint8_t* d, int64_t n, HostImage& im) {
  std::memset(&im.dev, 0, sizeof(im.dev));
  if (n < 4 || d[0] != 0xFF || d[1] != 0xD8) { pg_set_error("unsupported: not a JPEG file (no SOI)"); return PG_ERR_UNSUPPORTED; }
  uint16_t qt[4][64];
  bool have_qt[4] = {false, false, false, false};
  bool have_huff[2][2] = {{false, false}, {false, false}};
  int comp_id[3] = {0, 0, 0}, comp_tq[3] = {0, 0, 0};
  int64_t pos = 2;
  int dri = 0;
  bool have_sof = false;
  PgjImage& g = im.dev;
  while (true) {
    if (pos + 4 > n || d[pos] != 0xFF) { pg_set_error("unsupported: JPEG marker expected at byte %lld", (long long)pos); return PG_ERR_UNSUPPORTED; }
    while (pos + 1 < n && d[pos + 1] == 0xFF) ++pos;
    const int m = d[pos + 1];
    pos += 2;
    if (m == 0x01 || (m >= 0xD0 && m <= 0xD7)) continue;
    if (pos + 2 > n) { pg_set_error("unsupported: truncated JPEG"); return PG_ERR_UNSUPPORTED; }
    const int len = (d[pos] << 8) | d[pos + 1];
    if (len < 2 || pos + len > n) { pg_set_error("unsupported: truncated JPEG segment"); return PG_ERR_UNSUPPORTED; }
    const uint8_t* s = d + pos + 2;
    const int sl = len - 2;
    if (m == 0xDB) {
      int k = 0;
      while (k < sl) {
        const int pq = s[k] >> 4, tq = s[k] & 15;
        if (tq > 3 || k + 1 + (pq ? 128 : 64) > sl) { pg_set_error("unsupported: bad DQT"); return PG_ERR_UNSUPPORTED; }
        for (int i = 0; i < 64; ++i) {
          const int v = pq ? ((s[k + 1 + 2 * i] << 8) | s[k + 2 + 2 * i]) : s[k + 1 + i];
          qt[tq][pgj_zz_host[i]] = (uint16_t)v;
        }
        have_qt[tq] = true;
        k += 1 + (pq ? 128 : 64);
      }
    } else if (m == 0xC0 || m == 0xC1) {
      if (sl < 6 || s[0] != 8) { pg_set_error("unsupported: JPEG sample precision %d", sl ? s[0] : 0); return PG_ERR_UNSUPPORTED; }
      g.height = (s[1] << 8) | s[2];
      g.width = (s[3] << 8) | s[4];
      g.n_comps = s[5];
      if ((g.n_comps != 1 && g.n_comps != 3) || sl < 6 + 3 * g.n_comps || g.width < 1 || g.height < 1) {
        pg_set_error("unsupported: JPEG with %d components", g.n_comps);
        return PG_ERR_UNSUPPORTED;
      }
      for (int i = 0; i < g.n_comps; ++i) {
        comp_id[i] = s[6 + 3 * i];
        g.comp_h[i] = s[7 + 3 * i] >> 4;
        g.comp_v[i] = s[7 + 3 * i] & 15;
        comp_tq[i] = s[8 + 3 * i];
        if (comp_tq[i] > 3 || g.comp_h[i] < 1 || g.comp_v[i] < 1 || g.comp_h[i] > 2 || g.comp_v[i] > 2) {
          pg_set_error("unsupported: JPEG sampling factors %dx%d", g.comp_h[i], g.comp_v[i]);
          return PG_ERR_UNSUPPORTED;
        }
      }
      have_sof = true;
    } else if (m == 0xC2 || m == 0xC3 || (m >= 0xC5 && m <= 0xC7) || (m >= 0xC9 && m <= 0xCB) || (m >= 0xCD && m <= 0xCF)) {
      pg_set_error("unsupported: JPEG process SOF%d (only baseline / extended sequential Huffman is decoded on the device)", m - 0xC0);
      return PG_ERR_UNSUPPORTED;
    } else if (m == 0xC4) {
      int k = 0;
      while (k < sl) {
        if (k + 17 > sl) { pg_set_error("unsupported: bad DHT"); return PG_ERR_UNSUPPORTED; }
        const int tc = s[k] >> 4, th = s[k] & 15;
        int cnt = 0;
        for (int i = 0; i < 16; ++i) cnt += s[k + 1 + i];
        if (tc > 1 || th > 1 || k + 17 + cnt > sl || cnt > 256) { pg_set_error("unsupported: Huffman table class %d id %d", tc, th); return PG_ERR_UNSUPPORTED; }
        if (build_huff(s + k + 1, s + k + 17, cnt, g.huff[tc][th]) != 0) { pg_set_error("unsupported: bad Huffman table"); return PG_ERR_UNSUPPORTED; }
        have_huff[tc][th] = true;
        k += 17 + cnt;
      }
    } else if (m == 0xE1) {
      const int o = exif_orientation(s, sl);
      if (o > 1 && o <= 8) {
        pg_set_error("unsupported: EXIF orientation %d (cv2.imread would turn the image; left to the host decoder)", o);
        return PG_ERR_UNSUPPORTED;
      }
    } else if (m == 0xDD) {
      if (sl < 2) { pg_set_error("unsupported: bad DRI"); return PG_ERR_UNSUPPORTED; }
      dri = (s[0] << 8) | s[1];
    } else if (m == 0xDA) {
      if (!have_sof) { pg_set_error("unsupported: SOS before SOF"); return PG_ERR_UNSUPPORTED; }
      const int ns = sl ? s[0] : 0;
      if (ns != g.n_comps || sl < 1 + 2 * ns + 3) { pg_set_error("unsupported: multi-scan JPEG"); return PG_ERR_UNSUPPORTED; }
      for (int i = 0; i < ns; ++i) {
        int c = -1;
        for (int j = 0; j < g.n_comps; ++j) if (comp_id[j] == s[1 + 2 * i]) c = j;
        if (c != i) { pg_set_error("unsupported: scan component order"); return PG_ERR_UNSUPPORTED; }
        g.comp_dc[c] = s[2 + 2 * i] >> 4;
        g.comp_ac[c] = s[2 + 2 * i] & 15;
        if (g.comp_dc[c] > 1 || g.comp_ac[c] > 1 || !have_huff[0][g.comp_dc[c]] || !have_huff[1][g.comp_ac[c]] || !have_qt[comp_tq[c]]) {
          pg_set_error("unsupported: scan refers to a missing table");
          return PG_ERR_UNSUPPORTED;
        }
        std::memcpy(g.qt[c], qt[comp_tq[c]], sizeof(qt[0]));
      }
      im.scan_begin = pos + len;
      int64_t e = n;
      for (int64_t q = n - 2; q >= im.scan_begin; --q)
        if (d[q] == 0xFF && d[q + 1] == 0xD9) { e = q; break; }
      im.scan_end = e;
      break;
    }
    pos += len;
  }
  // geometry (a one-component scan is never interleaved: its MCU is one block)
  int hmax = 1, vmax = 1;
  if (g.n_comps == 1) {
    g.comp_h[0] = g.comp_v[0] = 1;
  } else {
    for (int i = 0; i < 3; ++i) { hmax = std::max(hmax, g.comp_h[i]); vmax = std::max(vmax, g.comp_v[i]); }
    if (g.comp_h[0] != hmax || g.comp_v[0] != vmax || g.comp_h[1] != 1 || g.comp_v[1] != 1 || g.comp_h[2] != 1 || g.comp_v[2] != 1) {
      pg_set_error("unsupported: JPEG subsampling layout (luma %dx%d)", g.comp_h[0], g.comp_v[0]);
      return PG_ERR_UNSUPPORTED;
    }
  }
  g.mcus_w = (g.width + 8 * hmax - 1) / (8 * hmax);
  g.mcus_h = (g.height + 8 * vmax - 1) / (8 * vmax);
  g.bpm = 0;
  int64_t off = 0;
  for (int c = 0; c < g.n_comps; ++c) {
    g.comp_bw[c] = g.mcus_w * g.comp_h[c];
    g.comp_bh[c] = g.mcus_h * g.comp_v[c];
    g.comp_coef_off[c] = off;
    off += (int64_t)g.comp_bw[c] * g.comp_bh[c] * 64;
    for (int y = 0; y < g.comp_v[c]; ++y)
      for (int x = 0; x < g.comp_h[c]; ++x) {
        g.blk_comp[g.bpm] = c; g.blk_dx[g.bpm] = x; g.blk_dy[g.bpm] = y;
        ++g.bpm;
      }
  }
  const int64_t mcus = (int64_t)g.mcus_w * g.mcus_h;
  if (mcus * g.bpm > 0x7fffffffll) { pg_set_error("unsupported: image too large"); return PG_ERR_UNSUPPORTED; }
  g.total_blocks = (int32_t)(mcus * g.bpm);
  g.restart_blocks = dri * g.bpm;
  im.n_intervals = dri ? (int32_t)((mcus + dri - 1) / dri) : 1;
  return PG_OK;
}

// ---- device records ---------------------------------------------------------------------------------------
struct ImgRec {            // per image, device
  int64_t src_off;         // entropy-coded segment inside the file blob
  int64_t src_len;
  int64_t cs_off;          // unstuffed stream inside the compact buffer (256-aligned)
  int64_t rst_off;         // first slot of the image's restart positions
  int32_t rst_cap;
  int32_t ub0, n_ub;       // unstuff blocks
  int64_t chunk0;          // first chunk
  int32_t n_chunks;
  int32_t cta0;            // first entropy CTA of the image (its CTAs are consecutive)
  int64_t coef_off;        // int16 elements
  int64_t plane_off;       // colour files: the three component planes (bytes into Scratch::planes)
  uint8_t* out;            // decoded page: one grey plane, or BGR interleaved
  int64_t pitch;
};

struct Scratch {           // device pointers into the caller's workspace
  PgjImage* img;
  ImgRec* rec;
  int32_t* cta_img;        // per entropy CTA: image
  int32_t* cta_chunk0;     // ... and its first chunk (image-local)
  int32_t* ub_img;         // per unstuff CTA: image
  int32_t* ub_kept; int32_t* ub_rst;   // per unstuff block counts -> exclusive offsets
  int64_t* img_len;        // unstuffed bytes per image
  int32_t* img_nrst;
  uint8_t* compact;
  int32_t* rst_pos;
  PgjChunkState* st[2];
  PgjEntry* ent;           // the entry state each chunk's stored exit was computed from
  PgjChunkState* cta_scan; // per entropy CTA: the segmented total of its chunks, then (in place) the state in front of it
  int32_t* counters;       // [MAX_ROUNDS + 2]: changes per round; [MAX_ROUNDS+1] = error flag
  int16_t* coef;
  uint8_t* planes;         // colour files: Y, Cb, Cr after the IDCT, each comp_bw*8 x comp_bh*8
};

__device__ __forceinline__ PgjStream stream_of(const Scratch& s, const ImgRec& r, int img) {
  PgjStream sv;
  sv.bytes = s.compact + r.cs_off;
  sv.n_bits = s.img_len[img] * 8;
  sv.rst_pos = s.rst_pos + r.rst_off;
  sv.n_rst = min(s.img_nrst[img], r.rst_cap);
  return sv;
}

// ---- D1-D3: unstuff ---------------------------------------------------------------------------------------
// Thread t of an unstuff block looks at 16 bytes of the file blob on an absolute 16-byte boundary (one LDG.128);
// the bytes before the segment's first / behind its last are masked out.  The neighbouring bytes a marker test
// needs come from the adjacent lanes.
__device__ __forceinline__ void unstuff_masks(const uint8_t* blob, const ImgRec& r, int64_t ub_local, uint32_t& keep,
                                              uint32_t& rst, uint4& bytes) {
  const int64_t abs0 = (r.src_off & ~(int64_t)15) + (ub_local * 256 + threadIdx.x) * 16;  // absolute, 16-aligned
  const int64_t rel0 = abs0 - r.src_off;                                                    // may be negative (< 16)
  const int lane = threadIdx.x & 31;
  const bool live = rel0 < r.src_len && rel0 > -16;
  bytes = live ? __ldg(reinterpret_cast<const uint4*>(blob + abs0)) : make_uint4(0u, 0u, 0u, 0u);
  uint32_t prev = __shfl_up_sync(0xffffffffu, bytes.w >> 24, 1);
  uint32_t next = __shfl_down_sync(0xffffffffu, bytes.x & 0xFFu, 1);
  if (lane == 0) prev = (live && abs0 > 0) ? blob[abs0 - 1] : 0u;
  if (lane == 31) next = (live && rel0 + 16 < r.src_len) ? blob[abs0 + 16] : 0u;
  keep = 0; rst = 0;
  if (!live) return;
  const uint32_t w[4] = {bytes.x, bytes.y, bytes.z, bytes.w};
  // bytes [lo, hi) of the sixteen belong to the segment; byte k + 1 exists inside it while k + 1 < lim
  const int lo = rel0 < 0 ? (int)-rel0 : 0;
  const int64_t left = r.src_len - rel0;
  const int hi = left < 16 ? (int)left : 16, lim = left < 17 ? (int)left : 17;
  pgj_unstuff_masks16(w, rel0 > 0 ? prev : 0u, next, lo, hi, lim, keep, rst);  // the byte before the segment's first does not count
}

__global__ void __launch_bounds__(256) jpeg_unstuff_count_kernel(const uint8_t* blob, Scratch s) {
  __shared__ int sm[34];
  const int img = s.ub_img[blockIdx.x];
  const ImgRec r = s.rec[img];
  uint32_t keep, rst;
  uint4 bytes;
  unstuff_masks(blob, r, blockIdx.x - r.ub0, keep, rst, bytes);
  int tk, tr;
  pg_block_exscan(__popc(keep), sm, &tk);
  pg_block_exscan(__popc(rst), sm, &tr);
  if (threadIdx.x == 0) { s.ub_kept[blockIdx.x] = tk; s.ub_rst[blockIdx.x] = tr; }
}

// one CTA per image: exclusive scan of its blocks' counts (in place), totals
__global__ void __launch_bounds__(1024) jpeg_unstuff_scan_kernel(Scratch s) {
  __shared__ int sm[34];
  const int img = blockIdx.x;
  const ImgRec r = s.rec[img];
  int base_k = 0, base_r = 0;
  for (int b0 = 0; b0 < r.n_ub; b0 += 1024) {
    const int b = b0 + threadIdx.x;
    const int vk = b < r.n_ub ? s.ub_kept[r.ub0 + b] : 0, vr = b < r.n_ub ? s.ub_rst[r.ub0 + b] : 0;
    int tk, tr;
    const int ek = pg_block_exscan(vk, sm, &tk);
    const int er = pg_block_exscan(vr, sm, &tr);
    if (b < r.n_ub) { s.ub_kept[r.ub0 + b] = base_k + ek; s.ub_rst[r.ub0 + b] = base_r + er; }
    base_k += tk; base_r += tr;
  }
  if (threadIdx.x == 0) { s.img_len[img] = base_k; s.img_nrst[img] = base_r; }
}

__global__ void __launch_bounds__(256) jpeg_unstuff_write_kernel(const uint8_t* blob, Scratch s) {
  __shared__ int sm[34];
  __shared__ __align__(16) uint8_t stage[UB_BYTES + 16];  // the block's kept bytes, packed; written out with coalesced stores
  const int img = s.ub_img[blockIdx.x];
  const ImgRec r = s.rec[img];
  uint32_t keep, rst;
  uint4 bytes;
  unstuff_masks(blob, r, blockIdx.x - r.ub0, keep, rst, bytes);
  int tk, tr;
  int ok = pg_block_exscan(__popc(keep), sm, &tk);
  int orr = pg_block_exscan(__popc(rst), sm, &tr) + s.ub_rst[blockIdx.x];
  const int out0 = s.ub_kept[blockIdx.x];
  const uint32_t w[4] = {bytes.x, bytes.y, bytes.z, bytes.w};
  for (uint32_t m = rst; m; m &= m - 1u) {  // rare: restart interval (orr + 1) starts at the next byte that is kept
    const int k = __ffs(m) - 1;
    if (orr < r.rst_cap) s.rst_pos[r.rst_off + orr] = out0 + ok + __popc(keep & ((1u << k) - 1u));
    ++orr;
  }
  PG_DEV_ASSERT(ok >= 0 && ok + __popc(keep) <= UB_BYTES && tk <= UB_BYTES);
  uint32_t sa = (uint32_t)__cvta_generic_to_shared(stage) + (uint32_t)ok;
  asm volatile("" : "+r"(sa));  // computed once: left alone, the compiler re-derives the shared window for every byte
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    if (keep >> k & 1u) {
      asm volatile("st.shared.u8 [%0], %1;" ::"r"(sa), "r"(w[k >> 2] >> (8 * (k & 3))) : "memory");
      ++sa;
    }
  }
  __syncthreads();
  // out: the bytes up to the first word boundary of the destination, whole words (each from two staged words), the rest
  uint8_t* dst = s.compact + r.cs_off + out0;
  const int head = min(tk, (int)((4u - (uint32_t)((uintptr_t)dst & 3u)) & 3u));
  const int nw = (tk - head) >> 2, tail0 = head + 4 * nw;
  if ((int)threadIdx.x < head) dst[threadIdx.x] = stage[threadIdx.x];
  const uint32_t* sw = reinterpret_cast<const uint32_t*>(stage);
  uint32_t* dw = reinterpret_cast<uint32_t*>(dst + head);
  for (int i = threadIdx.x; i < nw; i += 256) dw[i] = __funnelshift_r(sw[i], sw[i + 1], 8 * head);
  if ((int)threadIdx.x < tk - tail0) dst[tail0 + threadIdx.x] = stage[tail0 + threadIdx.x];
  // slack behind the stream: ones, so that nothing read there looks like a code word
  if (blockIdx.x == r.ub0) {
    uint8_t* tail = s.compact + r.cs_off + s.img_len[img];
    for (int k = threadIdx.x; k < STREAM_PAD; k += 256) tail[k] = 0xFF;
  }
}

// ---- D4: speculative pass ---------------------------------------------------------------------------------
__device__ __forceinline__ void load_image(PgjImage* dst, const PgjImage* src) {
  const int4* a = reinterpret_cast<const int4*>(src);
  int4* b = reinterpret_cast<int4*>(dst);
  for (int i = threadIdx.x; i < (int)(sizeof(PgjImage) / 16); i += blockDim.x) b[i] = a[i];
  __syncthreads();
}

// TABLES_SMEM: the image's tables are staged in shared memory (default) or read where they lie (global memory through
// L1; knob PG_JPEG_TABLES=global) — measured, see DESIGN.md.
template <bool TABLES_SMEM>
__global__ void __launch_bounds__(CHUNK_THREADS) jpeg_spec_kernel(Scratch s, int chunk_bits, int overlap_bits) {
  __shared__ __align__(16) uint8_t im_buf[TABLES_SMEM ? sizeof(PgjImage) : 16];
  const int img = s.cta_img[blockIdx.x];
  if (TABLES_SMEM) load_image(reinterpret_cast<PgjImage*>(im_buf), s.img + img);
  const PgjImage& im = TABLES_SMEM ? *reinterpret_cast<const PgjImage*>(im_buf) : s.img[img];
  const ImgRec r = s.rec[img];
  const int j = s.cta_chunk0[blockIdx.x] + threadIdx.x;
  if (j >= r.n_chunks) return;
  const PgjStream sv = stream_of(s, r, img);
  PgjChunkState out;
  PgjEntry en;
  pgj_spec_chunk(sv, im, j, chunk_bits, overlap_bits, en, out);
  s.st[0][r.chunk0 + j] = out;
  s.ent[r.chunk0 + j] = en;
}

// ---- D5: one sync round -----------------------------------------------------------------------------------
__global__ void __launch_bounds__(CHUNK_THREADS) jpeg_sync_kernel(Scratch s, int chunk_bits, int round) {
  __shared__ __align__(16) PgjImage im;
  const int img = s.cta_img[blockIdx.x];
  const ImgRec r = s.rec[img];
  const int j = s.cta_chunk0[blockIdx.x] + threadIdx.x;
  const PgjChunkState* in = s.st[(round - 1) & 1];
  PgjChunkState* outv = s.st[round & 1];
  // a round after one in which nothing was decoded again only carries the states over (uniform per launch)
  const bool idle = round > 1 && s.counters[round - 1] == 0;
  if (!idle) load_image(&im, s.img + img);
  if (j >= r.n_chunks) return;
  const int64_t g = r.chunk0 + j;
  PgjChunkState mine = in[g];
  if (!idle && j > 0 && mine.p != -2) {
    const PgjChunkState prev = in[g - 1];
    PgjEntry en = s.ent[g];
    if (prev.p != en.p || prev.c != en.c) {  // the exit stored for this chunk was reached from another entry
      const PgjStream sv = stream_of(s, r, img);
      pgj_sync_chunk(sv, im, j, chunk_bits, prev, en, mine);
      s.ent[g] = en;
      atomicAdd(&s.counters[round], 1);
    }
  }
  outv[g] = mine;
}

// ---- D6: decoder state at every chunk's entry ---------------------------------------------------------------
// A segmented exclusive scan over the chunks of an image: block count and DC-difference sums accumulate, a chunk
// that passed a restart boundary (anchor >= 0) starts a new segment.  Three steps: (a) every entropy CTA folds its
// 256 chunks into one value; (b) one CTA per image scans those values (a few hundred) in place; (c) the store
// kernel redoes the scan inside its CTA and adds what (b) left in front of it — the per-chunk result never
// touches memory.
struct ScanVal { int anchor, n, d0, d1, d2; };
__device__ __forceinline__ ScanVal scan_op(const ScanVal& a, const ScanVal& b) {  // a then b
  if (b.anchor >= 0) return b;
  return ScanVal{a.anchor, a.n + b.n, a.d0 + b.d0, a.d1 + b.d1, a.d2 + b.d2};
}
__device__ __forceinline__ ScanVal shfl_up(const ScanVal& v, int d) {
  return ScanVal{__shfl_up_sync(0xffffffffu, v.anchor, d), __shfl_up_sync(0xffffffffu, v.n, d),
                 __shfl_up_sync(0xffffffffu, v.d0, d), __shfl_up_sync(0xffffffffu, v.d1, d),
                 __shfl_up_sync(0xffffffffu, v.d2, d)};
}
__device__ __forceinline__ ScanVal scan_of(const PgjChunkState& cs) {
  if (cs.p == -2) return ScanVal{-1, 0, 0, 0, 0};
  return ScanVal{cs.anchor, cs.n, cs.dc[0], cs.dc[1], cs.dc[2]};
}
// Block-wide segmented scan (blockDim.x a multiple of 32, <= 1024): returns what lies in front of this thread inside
// the block; *total = the whole block.  `sm` needs 33 ScanVal.
__device__ __forceinline__ ScanVal block_scan(const ScanVal& v, ScanVal* sm, ScanVal* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const ScanVal ident{-1, 0, 0, 0, 0};
  ScanVal inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const ScanVal t = shfl_up(inc, d);
    if (lane >= d) inc = scan_op(t, inc);
  }
  if (lane == 31) sm[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    ScanVal w = lane < nw ? sm[lane] : ident;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const ScanVal t = shfl_up(w, d);
      if (lane >= d) w = scan_op(t, w);
    }
    if (lane < nw) sm[lane] = w;  // inclusive over warps
    if (lane == 31) sm[32] = w;   // (lanes >= nw hold the last warp's inclusive value)
  }
  __syncthreads();
  ScanVal before = warp > 0 ? sm[warp - 1] : ident;
  const ScanVal prev = shfl_up(inc, 1);
  if (lane > 0) before = scan_op(before, prev);
  *total = sm[32];
  __syncthreads();
  return before;
}

__global__ void __launch_bounds__(CHUNK_THREADS) jpeg_cta_total_kernel(Scratch s, int final_parity) {
  __shared__ ScanVal sm[33];
  const int img = s.cta_img[blockIdx.x];
  const ImgRec r = s.rec[img];
  const int j = s.cta_chunk0[blockIdx.x] + threadIdx.x;
  ScanVal v{-1, 0, 0, 0, 0};
  if (j < r.n_chunks) v = scan_of(s.st[final_parity][r.chunk0 + j]);
  ScanVal total;
  block_scan(v, sm, &total);
  if (threadIdx.x == 0) {
    PgjChunkState o;
    o.p = 0; o.c = 0; o.anchor = total.anchor; o.n = total.n; o.dc[0] = total.d0; o.dc[1] = total.d1; o.dc[2] = total.d2;
    s.cta_scan[blockIdx.x] = o;
  }
}

__global__ void __launch_bounds__(256) jpeg_cta_carry_kernel(Scratch s) {
  __shared__ ScanVal sm[33];
  __shared__ ScanVal carry_sm;
  const int img = blockIdx.x;
  const ImgRec r = s.rec[img];
  const int n_cta = (r.n_chunks + CHUNK_THREADS - 1) / CHUNK_THREADS;
  if (threadIdx.x == 0) carry_sm = ScanVal{0, 0, 0, 0, 0};  // chunk 0 enters restart interval 0 at block 0
  __syncthreads();
  for (int b0 = 0; b0 < n_cta; b0 += 256) {
    const int b = b0 + threadIdx.x;
    ScanVal v{-1, 0, 0, 0, 0};
    if (b < n_cta) {
      const PgjChunkState cs = s.cta_scan[r.cta0 + b];
      v = ScanVal{cs.anchor, cs.n, cs.dc[0], cs.dc[1], cs.dc[2]};
    }
    ScanVal total;
    const ScanVal before = block_scan(v, sm, &total);
    const ScanVal carry = carry_sm;
    if (b < n_cta) {
      const ScanVal e = scan_op(carry, before);
      PgjChunkState o;
      o.p = 0; o.c = 0; o.anchor = e.anchor; o.n = e.n; o.dc[0] = e.d0; o.dc[1] = e.d1; o.dc[2] = e.d2;
      s.cta_scan[r.cta0 + b] = o;
    }
    __syncthreads();
    if (threadIdx.x == 0) carry_sm = scan_op(carry, total);
    __syncthreads();
  }
}

// ---- D7: coefficient store --------------------------------------------------------------------------------
// Every block is assembled in the thread's own 128 bytes of shared memory and leaves as one full line (eight
// 16-byte stores): no zero-fill pass over the coefficient buffer, no two-byte scatter into global memory.
constexpr int BLK_PITCH = 72;  // int16 per thread (64 + 8): 16-byte accesses of a quarter warp fall on distinct banks
struct SmemBlockSink {
  int16_t* sm;      // this thread's block
  int16_t* base;    // the image's coefficient base
  int pred;
  int64_t cap;      // coefficients the image owns (checked build)
  __device__ __forceinline__ void begin(int p) {
    pred = p;
    uint4* z = reinterpret_cast<uint4*>(sm);
#pragma unroll
    for (int k = 0; k < 8; ++k) z[k] = make_uint4(0u, 0u, 0u, 0u);
  }
  __device__ __forceinline__ void dc(int diff) { sm[0] = (int16_t)(pred + diff); }
  __device__ __forceinline__ void ac(int idx, int v) {
    PG_DEV_ASSERT(idx >= 1 && idx < 64);
    sm[idx] = (int16_t)v;
  }
  __device__ __forceinline__ void end(int64_t ci) {
    PG_DEV_ASSERT(ci >= 0 && ci + 64 <= cap && (ci & 63) == 0);
    const uint4* src = reinterpret_cast<const uint4*>(sm);
    uint4* dst = reinterpret_cast<uint4*>(base + ci);
#pragma unroll
    for (int k = 0; k < 8; ++k) dst[k] = src[k];
  }
};

template <bool TABLES_SMEM>
__global__ void __launch_bounds__(CHUNK_THREADS) jpeg_store_kernel(Scratch s, int chunk_bits, int final_parity) {
  __shared__ __align__(16) uint8_t im_buf[TABLES_SMEM ? sizeof(PgjImage) : 16];
  __shared__ ScanVal scan_sm[33];
  extern __shared__ __align__(16) int16_t blocks[];  // CHUNK_THREADS * BLK_PITCH (dynamic: with the tables > 48 KB)
  const int img = s.cta_img[blockIdx.x];
  if (TABLES_SMEM) load_image(reinterpret_cast<PgjImage*>(im_buf), s.img + img);
  const PgjImage& im = TABLES_SMEM ? *reinterpret_cast<const PgjImage*>(im_buf) : s.img[img];
  const ImgRec r = s.rec[img];
  const int j = s.cta_chunk0[blockIdx.x] + threadIdx.x;
  const PgjChunkState* st = s.st[final_parity];
  // D6 (c): the state in front of this chunk = what lies in front of the CTA + the chunks before it inside the CTA
  ScanVal v{-1, 0, 0, 0, 0};
  if (j < r.n_chunks) v = scan_of(st[r.chunk0 + j]);
  ScanVal total;
  ScanVal before = block_scan(v, scan_sm, &total);
  {
    const PgjChunkState cs = s.cta_scan[blockIdx.x];
    before = scan_op(ScanVal{cs.anchor, cs.n, cs.dc[0], cs.dc[1], cs.dc[2]}, before);
  }
  if (j >= r.n_chunks) return;
  const PgjStream sv = stream_of(s, r, img);
  const int64_t b0 = (int64_t)j * chunk_bits, b1 = b0 + chunk_bits;
  if (b0 >= sv.n_bits) return;
  const int64_t ep = j == 0 ? 0 : st[r.chunk0 + j - 1].p;
  const int ec = j == 0 ? 0 : st[r.chunk0 + j - 1].c;
  if (ep < 0 || ep >= b1) return;  // no block starts inside this chunk
  const int blk = max(before.anchor, 0) * im.restart_blocks + before.n;
  int64_t cap = 0;
  for (int c = 0; c < im.n_comps; ++c) cap += (int64_t)im.comp_bw[c] * im.comp_bh[c] * 64;
  PG_DEV_ASSERT(blk >= 0 && ep >= 0 && ep <= sv.n_bits);
  SmemBlockSink sink{blocks + threadIdx.x * BLK_PITCH, s.coef + r.coef_off, 0, cap};
  pgj_span_store(sv, im, ep, ec, b1, blk, before.d0, before.d1, before.d2, sink);
}

// ---- D8: inverse DCT, one thread per block -------------------------------------------------------------------
// Greyscale files: straight into the page.  Colour files (blockIdx.z = component): into the component's own plane,
// whole blocks; D9 makes the page from the three planes.
__global__ void __launch_bounds__(128) jpeg_idct_kernel(Scratch s) {
  const int img = blockIdx.y, comp = blockIdx.z;
  const ImgRec r = s.rec[img];
  const PgjImage* im = s.img + img;
  if (comp >= im->n_comps) return;
  const int bw = im->comp_bw[comp], bh = im->comp_bh[comp];
  const int64_t nb = (int64_t)bw * bh;
  const bool grey = im->n_comps == 1;
  const int w = grey ? im->width : bw * 8, h = grey ? im->height : bh * 8;
  uint8_t* plane = r.out;
  int64_t pitch = r.pitch;
  if (!grey) {
    int64_t off = r.plane_off;
    for (int c = 0; c < comp; ++c) off += (int64_t)im->comp_bw[c] * im->comp_bh[c] * 64;
    plane = s.planes + off;
    pitch = bw * 8;
  }
  __shared__ uint16_t q[64];
  if (threadIdx.x < 64) q[threadIdx.x] = im->qt[comp][threadIdx.x];
  __syncthreads();
  for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < nb; b += (int64_t)gridDim.x * blockDim.x) {
    const int by = (int)(b / bw), bx = (int)(b - (int64_t)by * bw);
    const int16_t* src = s.coef + r.coef_off + im->comp_coef_off[comp] + b * 64;
    __align__(16) int16_t c[64];
    const int4* v = reinterpret_cast<const int4*>(src);
#pragma unroll
    for (int k = 0; k < 8; ++k) reinterpret_cast<int4*>(c)[k] = __ldg(v + k);
    __align__(8) uint8_t px[64];
    pgj_idct_block(c, q, px);
    const int x0 = bx * 8, y0 = by * 8;
    if (x0 >= w) continue;
#pragma unroll
    for (int y = 0; y < 8; ++y) {
      if (y0 + y >= h) break;
      uint8_t* dst = plane + (int64_t)(y0 + y) * pitch + x0;
      PG_DEV_ASSERT(x0 + (x0 + 8 <= w ? 8 : w - x0) <= pitch);
      if (x0 + 8 <= w) {
        *reinterpret_cast<uint2*>(dst) = *reinterpret_cast<const uint2*>(px + 8 * y);
      } else {
        for (int x = 0; x0 + x < w; ++x) dst[x] = px[8 * y + x];
      }
    }
  }
}

// ---- D9: colour files — fancy chroma upsampling + YCbCr -> BGR, four pixels (12 bytes) per thread --------------------
__global__ void __launch_bounds__(256) jpeg_colour_kernel(Scratch s) {
  const int img = blockIdx.y;
  const PgjImage* im = s.img + img;
  if (im->n_comps != 3) return;
  const ImgRec r = s.rec[img];
  const int w = im->width, h = im->height;
  const int hmax = im->comp_h[0], vmax = im->comp_v[0];
  const int p0 = im->comp_bw[0] * 8, p1 = im->comp_bw[1] * 8, p2 = im->comp_bw[2] * 8;
  const uint8_t* yp = s.planes + r.plane_off;
  const uint8_t* cbp = yp + (int64_t)im->comp_bw[0] * im->comp_bh[0] * 64;
  const uint8_t* crp = cbp + (int64_t)im->comp_bw[1] * im->comp_bh[1] * 64;
  const int fx = hmax / im->comp_h[1], fy = vmax / im->comp_v[1];
  const int dw = (w * im->comp_h[1] + hmax - 1) / hmax, dh = (h * im->comp_v[1] + vmax - 1) / vmax;
  const int quads = (w + 3) >> 2;
  const int64_t total = (int64_t)quads * h;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int y = (int)(t / quads), x0 = (int)(t - (int64_t)y * quads) * 4;
    uint8_t px[12];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int x = x0 + k < w ? x0 + k : w - 1;
      const int yy = yp[(int64_t)y * p0 + x];
      const int cb = pgj_upsample(cbp, p1, dw, dh, fx, fy, x, y), cr = pgj_upsample(crp, p2, dw, dh, fx, fy, x, y);
      pgj_ycc_to_bgr(yy, cb, cr, px[3 * k], px[3 * k + 1], px[3 * k + 2]);
    }
    uint8_t* dst = r.out + (int64_t)y * r.pitch + 3 * x0;
    if (x0 + 4 <= w) {
      uint32_t* d32 = reinterpret_cast<uint32_t*>(dst);  // 12 * (x0 / 4) bytes into a 16-byte aligned row
      d32[0] = px[0] | px[1] << 8 | px[2] << 16 | (uint32_t)px[3] << 24;
      d32[1] = px[4] | px[5] << 8 | px[6] << 16 | (uint32_t)px[7] << 24;
      d32[2] = px[8] | px[9] << 8 | px[10] << 16 | (uint32_t)px[11] << 24;
    } else {
      for (int k = 0; k < 3 * (w - x0); ++k) dst[k] = px[k];
    }
  }
}

}  // namespace

// ================================================================================================================
struct PgJpegDecoder {
  std::vector<HostImage> images;
  std::vector<int64_t> file_off;
  int32_t chunk_bytes = 512;
  int32_t rounds = 3;
  // layout computed by set_files
  std::vector<ImgRec> rec;
  std::vector<int32_t> cta_img, cta_chunk0, ub_img;
  int64_t total_chunks = 0, total_ub = 0, total_rst = 0, compact_bytes = 0, coef_elems = 0, plane_bytes = 0;
  bool any_colour = false;
  size_t ws_bytes = 0;
  // pinned staging for the per-call tables (ring, so that a call may be prepared while the previous one's copy runs)
  struct Stage { void* host = nullptr; size_t cap = 0; cudaEvent_t ev = nullptr; };
  Stage stage[4];
  int next_stage = 0;
  void* staged_ws = nullptr;  // workspace pg_jpeg_stage_tables prepared for the current files (consumed by the decode)
  Scratch last{};  // device view of the last decode (status read-back)
};

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

extern "C" int pg_jpeg_decoder_create(PgJpegDecoder** out) {
  PG_REQUIRE(out != nullptr, "decoder");
  *out = new PgJpegDecoder();
  return PG_OK;
}

extern "C" void pg_jpeg_decoder_destroy(PgJpegDecoder* d) {
  if (!d) return;
  for (auto& st : d->stage) {
    if (st.host) cudaFreeHost(st.host);
    if (st.ev) cudaEventDestroy(st.ev);
  }
  delete d;
}

extern "C" int pg_jpeg_decoder_configure(PgJpegDecoder* d, int32_t chunk_bytes, int32_t sync_rounds) {
  PG_REQUIRE(d != nullptr, "decoder");
  PG_REQUIRE(chunk_bytes >= 64 && chunk_bytes <= 65536 && (chunk_bytes & (chunk_bytes - 1)) == 0, "chunk_bytes: power of two in [64, 65536]");
  PG_REQUIRE(sync_rounds >= 1 && sync_rounds <= MAX_ROUNDS, "sync_rounds");
  d->chunk_bytes = chunk_bytes;
  d->rounds = sync_rounds;
  return PG_OK;
}

extern "C" int pg_jpeg_decoder_set_files(PgJpegDecoder* d, const uint8_t* blob, const int64_t* file_off, int32_t n) {
  PG_REQUIRE(d && blob && file_off && n > 0, "set_files arguments");
  d->images.assign((size_t)n, HostImage());
  d->staged_ws = nullptr;
  d->file_off.assign(file_off, file_off + n + 1);
  d->rec.assign((size_t)n, ImgRec());
  d->cta_img.clear(); d->cta_chunk0.clear(); d->ub_img.clear();
  int64_t cs = 0, rst = 0, chunks = 0, ub = 0, coef = 0, planes = 0;
  d->any_colour = false;
  for (int i = 0; i < n; ++i) {
    PG_REQUIRE(file_off[i + 1] > file_off[i], "file offsets must increase");
    const int rc = parse_jpeg(blob + file_off[i], file_off[i + 1] - file_off[i], d->images[(size_t)i]);
    if (rc != PG_OK) return rc;
    HostImage& im = d->images[(size_t)i];
    ImgRec& r = d->rec[(size_t)i];
    r.src_off = file_off[i] + im.scan_begin;
    r.src_len = im.scan_end - im.scan_begin;
    if (r.src_len < 1 || r.src_len > 0x7fff0000ll) { pg_set_error("unsupported: entropy-coded segment of %lld bytes", (long long)r.src_len); return PG_ERR_UNSUPPORTED; }
    r.cs_off = cs;
    cs += (int64_t)align_up((size_t)r.src_len + STREAM_PAD, 256);
    r.rst_off = rst;
    r.rst_cap = im.n_intervals;  // restart k+1 for k < n_intervals - 1; one spare slot
    rst += r.rst_cap;
    r.ub0 = (int32_t)ub;
    r.n_ub = (int32_t)(((r.src_off & 15) + r.src_len + UB_BYTES - 1) / UB_BYTES);  // blocks start on a 16-byte boundary
    for (int b = 0; b < r.n_ub; ++b) d->ub_img.push_back(i);
    ub += r.n_ub;
    r.chunk0 = chunks;
    r.n_chunks = (int32_t)((r.src_len + d->chunk_bytes - 1) / d->chunk_bytes);
    r.cta0 = (int32_t)d->cta_img.size();
    for (int c0 = 0; c0 < r.n_chunks; c0 += CHUNK_THREADS) { d->cta_img.push_back(i); d->cta_chunk0.push_back(c0); }
    chunks += r.n_chunks;
    r.coef_off = coef;
    coef += (int64_t)im.dev.total_blocks * 64;
    r.plane_off = planes;
    if (im.dev.n_comps == 3) {
      planes += (int64_t)im.dev.total_blocks * 64;  // one byte per coefficient: the three planes after the IDCT
      planes = (int64_t)align_up((size_t)planes, 256);
      d->any_colour = true;
    }
    r.out = nullptr; r.pitch = 0;
  }
  d->total_chunks = chunks; d->total_ub = ub; d->total_rst = rst; d->compact_bytes = cs; d->coef_elems = coef;
  d->plane_bytes = planes;
  // workspace layout (sizes only; pointers are formed at decode time)
  size_t w = 0;
  auto add = [&](size_t bytes) { w = align_up(w, 256) + bytes; };
  add((size_t)n * sizeof(PgjImage)); add((size_t)n * sizeof(ImgRec));
  add(d->cta_img.size() * 4); add(d->cta_chunk0.size() * 4); add(d->ub_img.size() * 4);
  add((size_t)ub * 4); add((size_t)ub * 4); add((size_t)n * 8); add((size_t)n * 4);
  add((size_t)cs); add((size_t)rst * 4 + 4);
  add((size_t)chunks * sizeof(PgjChunkState)); add((size_t)chunks * sizeof(PgjChunkState));
  add((size_t)chunks * sizeof(PgjEntry));
  add(d->cta_img.size() * sizeof(PgjChunkState));
  add((MAX_ROUNDS + 2) * 4);
  add((size_t)coef * 2);
  add((size_t)planes);
  d->ws_bytes = align_up(w, 256) + 256;
  return PG_OK;
}

extern "C" int pg_jpeg_decoder_image_info(const PgJpegDecoder* d, int32_t i, int32_t* w, int32_t* h, int32_t* channels) {
  PG_REQUIRE(d && i >= 0 && i < (int32_t)d->images.size(), "image index");
  if (w) *w = d->images[(size_t)i].dev.width;
  if (h) *h = d->images[(size_t)i].dev.height;
  if (channels) *channels = d->images[(size_t)i].dev.n_comps;
  return PG_OK;
}

extern "C" int64_t pg_jpeg_workspace_bytes(const PgJpegDecoder* d) { return d ? (int64_t)d->ws_bytes : 0; }

// Validates the outputs, carves the workspace and uploads the batch's tables on `s`.
static int stage_tables(PgJpegDecoder* d, uint8_t* const* out_ptrs, const int64_t* pitches, void* workspace,
                        int64_t workspace_bytes, cudaStream_t s) {
  const int n = (int)d->images.size();
  PG_REQUIRE(n > 0, "pg_jpeg_decoder_set_files first");
  if ((size_t)workspace_bytes < d->ws_bytes) { pg_set_error("workspace too small: %lld < %zu", (long long)workspace_bytes, d->ws_bytes); return PG_ERR_WORKSPACE; }
  PG_REQUIRE(((uintptr_t)workspace & 255) == 0, "workspace must be 256-byte aligned");
  for (int i = 0; i < n; ++i) {
    const PgjImage& g = d->images[(size_t)i].dev;
    PG_REQUIRE(out_ptrs[i] != nullptr && pitches[i] >= (int64_t)g.width * (g.n_comps == 1 ? 1 : 3), "output pointer / pitch");
    PG_REQUIRE(((uintptr_t)out_ptrs[i] & 15) == 0 && pitches[i] % 16 == 0, "output must be 16-byte aligned with pitch % 16 == 0");
    d->rec[(size_t)i].out = out_ptrs[i];
    d->rec[(size_t)i].pitch = pitches[i];
  }
  // carve the workspace
  uint8_t* base = reinterpret_cast<uint8_t*>(workspace);
  size_t w = 0;
  auto take = [&](size_t bytes) { w = align_up(w, 256); uint8_t* p = base + w; w += bytes; return p; };
  Scratch sc;
  sc.img = reinterpret_cast<PgjImage*>(take((size_t)n * sizeof(PgjImage)));
  sc.rec = reinterpret_cast<ImgRec*>(take((size_t)n * sizeof(ImgRec)));
  sc.cta_img = reinterpret_cast<int32_t*>(take(d->cta_img.size() * 4));
  sc.cta_chunk0 = reinterpret_cast<int32_t*>(take(d->cta_chunk0.size() * 4));
  sc.ub_img = reinterpret_cast<int32_t*>(take(d->ub_img.size() * 4));
  sc.ub_kept = reinterpret_cast<int32_t*>(take((size_t)d->total_ub * 4));
  sc.ub_rst = reinterpret_cast<int32_t*>(take((size_t)d->total_ub * 4));
  sc.img_len = reinterpret_cast<int64_t*>(take((size_t)n * 8));
  sc.img_nrst = reinterpret_cast<int32_t*>(take((size_t)n * 4));
  sc.compact = take((size_t)d->compact_bytes);
  sc.rst_pos = reinterpret_cast<int32_t*>(take((size_t)d->total_rst * 4 + 4));
  sc.st[0] = reinterpret_cast<PgjChunkState*>(take((size_t)d->total_chunks * sizeof(PgjChunkState)));
  sc.st[1] = reinterpret_cast<PgjChunkState*>(take((size_t)d->total_chunks * sizeof(PgjChunkState)));
  sc.ent = reinterpret_cast<PgjEntry*>(take((size_t)d->total_chunks * sizeof(PgjEntry)));
  sc.cta_scan = reinterpret_cast<PgjChunkState*>(take(d->cta_img.size() * sizeof(PgjChunkState)));
  sc.counters = reinterpret_cast<int32_t*>(take((MAX_ROUNDS + 2) * 4));
  sc.coef = reinterpret_cast<int16_t*>(take((size_t)d->coef_elems * 2));
  sc.planes = take((size_t)d->plane_bytes);
  d->last = sc;

  // per-call tables: one pinned staging block
  const size_t sz_img = (size_t)n * sizeof(PgjImage), sz_rec = (size_t)n * sizeof(ImgRec);
  const size_t sz_cta = d->cta_img.size() * 4, sz_ub = d->ub_img.size() * 4;
  const size_t o_img = 0, o_rec = align_up(o_img + sz_img, 256), o_ci = align_up(o_rec + sz_rec, 256),
               o_cc = align_up(o_ci + sz_cta, 256), o_ub = align_up(o_cc + sz_cta, 256), total = align_up(o_ub + sz_ub, 256);
  PgJpegDecoder::Stage& st = d->stage[d->next_stage];
  d->next_stage = (d->next_stage + 1) % 4;
  if (!st.ev) PG_CUDA_TRY(cudaEventCreateWithFlags(&st.ev, cudaEventDisableTiming));
  else PG_CUDA_TRY(cudaEventSynchronize(st.ev));
  if (st.cap < total) {
    // (re)size the whole ring at once: a pinned allocation synchronises the device, so it must not recur in a
    // steady stream of calls
    const size_t cap = std::max(total * 2, (size_t)1 << 20);
    for (auto& other : d->stage) {
      if (other.ev) PG_CUDA_TRY(cudaEventSynchronize(other.ev));
      if (other.host) cudaFreeHost(other.host);
      other.host = nullptr;
      other.cap = 0;
      PG_CUDA_TRY(cudaHostAlloc(&other.host, cap, cudaHostAllocDefault));
      other.cap = cap;
    }
  }
  uint8_t* hs = reinterpret_cast<uint8_t*>(st.host);
  for (int i = 0; i < n; ++i) std::memcpy(hs + o_img + (size_t)i * sizeof(PgjImage), &d->images[(size_t)i].dev, sizeof(PgjImage));
  std::memcpy(hs + o_rec, d->rec.data(), sz_rec);
  std::memcpy(hs + o_ci, d->cta_img.data(), sz_cta);
  std::memcpy(hs + o_cc, d->cta_chunk0.data(), sz_cta);
  std::memcpy(hs + o_ub, d->ub_img.data(), sz_ub);
  PG_CUDA_TRY(cudaMemcpyAsync(sc.img, hs + o_img, sz_img, cudaMemcpyHostToDevice, s));
  PG_CUDA_TRY(cudaMemcpyAsync(sc.rec, hs + o_rec, sz_rec, cudaMemcpyHostToDevice, s));
  PG_CUDA_TRY(cudaMemcpyAsync(sc.cta_img, hs + o_ci, sz_cta, cudaMemcpyHostToDevice, s));
  PG_CUDA_TRY(cudaMemcpyAsync(sc.cta_chunk0, hs + o_cc, sz_cta, cudaMemcpyHostToDevice, s));
  PG_CUDA_TRY(cudaMemcpyAsync(sc.ub_img, hs + o_ub, sz_ub, cudaMemcpyHostToDevice, s));
  PG_CUDA_TRY(cudaEventRecord(st.ev, s));
  d->staged_ws = workspace;
  return PG_OK;
}

extern "C" int pg_jpeg_stage_tables(PgJpegDecoder* d, uint8_t* const* out_ptrs, const int64_t* pitches, void* workspace,
                                    int64_t workspace_bytes, void* copy_stream) {
  PG_REQUIRE(d && out_ptrs && pitches && workspace, "stage arguments");
  return stage_tables(d, out_ptrs, pitches, workspace, workspace_bytes, (cudaStream_t)copy_stream);
}

extern "C" int pg_jpeg_decode(PgJpegDecoder* d, const uint8_t* blob_dev, uint8_t* const* out_ptrs, const int64_t* pitches,
                              void* workspace, int64_t workspace_bytes, void* stream) {
  PG_REQUIRE(d && blob_dev && out_ptrs && pitches && workspace, "decode arguments");
  PG_REQUIRE(((uintptr_t)blob_dev & 15) == 0, "blob_dev must be 16-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  if (d->staged_ws != workspace) {  // tables not uploaded by pg_jpeg_stage_tables for this batch: do it here
    const int rc = stage_tables(d, out_ptrs, pitches, workspace, workspace_bytes, s);
    if (rc != PG_OK) return rc;
  }
  d->staged_ws = nullptr;  // consumed
  const int n = (int)d->images.size();
  const Scratch sc = d->last;
  PG_CUDA_TRY(cudaMemsetAsync(sc.counters, 0, (MAX_ROUNDS + 2) * 4, s));

  const int chunk_bits = d->chunk_bytes * 8;
  const unsigned n_ub = (unsigned)d->total_ub, n_cta = (unsigned)d->cta_img.size();
  jpeg_unstuff_count_kernel<<<n_ub, 256, 0, s>>>(blob_dev, sc);
  jpeg_unstuff_scan_kernel<<<(unsigned)n, 1024, 0, s>>>(sc);
  jpeg_unstuff_write_kernel<<<n_ub, 256, 0, s>>>(blob_dev, sc);
  bool tables_smem = true;
  if (const char* e = getenv("PG_JPEG_TABLES")) tables_smem = std::strcmp(e, "global") != 0;  // tuning knob
  if (tables_smem) jpeg_spec_kernel<true><<<n_cta, CHUNK_THREADS, 0, s>>>(sc, chunk_bits, pgj_overlap_bits(chunk_bits));
  else jpeg_spec_kernel<false><<<n_cta, CHUNK_THREADS, 0, s>>>(sc, chunk_bits, pgj_overlap_bits(chunk_bits));
  for (int r = 1; r <= d->rounds; ++r) jpeg_sync_kernel<<<n_cta, CHUNK_THREADS, 0, s>>>(sc, chunk_bits, r);
  jpeg_cta_total_kernel<<<n_cta, CHUNK_THREADS, 0, s>>>(sc, d->rounds & 1);
  jpeg_cta_carry_kernel<<<(unsigned)n, 256, 0, s>>>(sc);
  constexpr size_t kStoreSmem = (size_t)CHUNK_THREADS * BLK_PITCH * sizeof(int16_t);
  if (tables_smem) {
    PG_CUDA_TRY(cudaFuncSetAttribute(jpeg_store_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kStoreSmem));
    jpeg_store_kernel<true><<<n_cta, CHUNK_THREADS, kStoreSmem, s>>>(sc, chunk_bits, d->rounds & 1);
  } else {
    PG_CUDA_TRY(cudaFuncSetAttribute(jpeg_store_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kStoreSmem));
    jpeg_store_kernel<false><<<n_cta, CHUNK_THREADS, kStoreSmem, s>>>(sc, chunk_bits, d->rounds & 1);
  }
  int max_blocks = 0;
  for (int i = 0; i < n; ++i) max_blocks = std::max(max_blocks, d->images[(size_t)i].dev.total_blocks);
  dim3 grid((unsigned)std::min(4096, (max_blocks + 127) / 128), (unsigned)n, d->any_colour ? 3u : 1u);
  jpeg_idct_kernel<<<grid, 128, 0, s>>>(sc);
  if (d->any_colour) {
    int64_t max_quads = 0;
    for (int i = 0; i < n; ++i) {
      const PgjImage& g = d->images[(size_t)i].dev;
      if (g.n_comps == 3) max_quads = std::max(max_quads, (int64_t)((g.width + 3) / 4) * g.height);
    }
    dim3 cgrid((unsigned)std::min<int64_t>(8192, (max_quads + 255) / 256), (unsigned)n);
    jpeg_colour_kernel<<<cgrid, 256, 0, s>>>(sc);
  }
  PG_LAUNCH_CHECK();
  return PG_OK;
}

// After the stream has been synchronised: stats[0] = PG_OK, or PG_ERR_UNSUPPORTED when the fixed number of sync
// rounds did not reach the fixed point (the caller raises `sync_rounds` and decodes again); stats[1] = rounds
// that changed something; stats[2] = states replaced in round 1; stats[3] = chunks.
extern "C" int pg_jpeg_decode_status(const PgJpegDecoder* d, int64_t stats[4]) {
  PG_REQUIRE(d && stats && d->last.counters, "decoder / stats");
  int32_t c[MAX_ROUNDS + 2];
  PG_CUDA_TRY(cudaMemcpy(c, d->last.counters, sizeof(c), cudaMemcpyDeviceToHost));
  int used = 0;
  for (int r = 1; r <= d->rounds; ++r) if (c[r]) used = r;
  stats[0] = c[d->rounds] != 0 ? PG_ERR_UNSUPPORTED : PG_OK;
  stats[1] = used;
  stats[2] = c[1];
  stats[3] = d->total_chunks;
  return PG_OK;
}

struct HostBlockSink {
  int16_t* base;
  int16_t blk[64];
  int pred;
  void begin(int p) { pred = p; std::memset(blk, 0, sizeof(blk)); }
  void dc(int diff) { blk[0] = (int16_t)(pred + diff); }
  void ac(int idx, int v) { blk[idx] = (int16_t)v; }
  void end(int64_t ci) { std::memcpy(base + ci, blk, sizeof(blk)); }
};

// ================================================================================================================
// Host evaluation of the same inline code, chunk by chunk in the kernels' order (CPU test-suite): decodes one
// greyscale file into out[height, pitch].  stats[0] = sync rounds that changed something, [1] = states replaced in
// round 1, [2] = chunks, [3] = restart markers found.
extern "C" int pg_hostcheck_jpeg_decode(const uint8_t* file, int64_t len, int32_t chunk_bytes, int32_t max_rounds,
                                        uint8_t* out, int64_t pitch, int32_t* width, int32_t* height, int64_t stats[4]) {
  PG_REQUIRE(file && len > 0 && chunk_bytes >= 8, "hostcheck arguments");
  HostImage hi;
  const int rc = parse_jpeg(file, len, hi);
  if (rc != PG_OK) return rc;
  const PgjImage& im = hi.dev;
  if (width) *width = im.width;
  if (height) *height = im.height;
  if (!out) {  // header only
    if (stats) { stats[0] = stats[1] = stats[2] = 0; stats[3] = im.n_comps; }
    return PG_OK;
  }
  PG_REQUIRE(pitch >= (int64_t)im.width * (im.n_comps == 1 ? 1 : 3), "pitch");
  // D1-D3
  const uint8_t* p = file + hi.scan_begin;
  const int64_t sl = hi.scan_end - hi.scan_begin;
  std::vector<uint8_t> compact;
  std::vector<int32_t> rst;
  compact.reserve((size_t)sl + STREAM_PAD + 8);
  for (int64_t j = 0; j < sl; ++j) {
    const uint8_t prev = j > 0 ? p[j - 1] : 0, cur = p[j], next = j + 1 < sl ? p[j + 1] : 0;
    if (pgj_rst_starts(cur, next)) rst.push_back((int32_t)compact.size());
    if (pgj_keep_byte(prev, cur, next)) compact.push_back(cur);
  }
  // the kernels' sixteen-bytes-at-once form of the two tests must say the same, at every alignment of the segment
  for (int64_t g0 = -(int64_t)(hi.scan_begin & 15); g0 < sl; g0 += 16) {
    uint32_t w[4] = {0u, 0u, 0u, 0u};
    for (int k = 0; k < 16; ++k)
      if (g0 + k >= -hi.scan_begin && hi.scan_begin + g0 + k < len) w[k >> 2] |= (uint32_t)p[g0 + k] << (8 * (k & 3));
    const int lo = g0 < 0 ? (int)-g0 : 0;
    const int64_t left = sl - g0;
    const int hi16 = left < 16 ? (int)left : 16, lim = left < 17 ? (int)left : 17;
    const uint32_t before = g0 > 0 ? p[g0 - 1] : 0u, after = hi.scan_begin + g0 + 16 < len ? p[g0 + 16] : 0u;
    uint32_t keep, rstm;
    pgj_unstuff_masks16(w, before, after, lo, hi16, lim, keep, rstm);
    for (int k = lo; k < hi16; ++k) {
      const int64_t j = g0 + k;
      const uint8_t pv = j > 0 ? p[j - 1] : 0, cur = p[j], nx = j + 1 < sl ? p[j + 1] : 0;
      PG_REQUIRE(((keep >> k) & 1u) == (pgj_keep_byte(pv, cur, nx) ? 1u : 0u) &&
                     ((rstm >> k) & 1u) == (pgj_rst_starts(cur, nx) ? 1u : 0u), "pgj_unstuff_masks16 disagrees with the byte tests");
    }
    PG_REQUIRE((keep >> hi16) == 0u && (rstm >> hi16) == 0u && (keep & ((1u << lo) - 1u)) == 0u, "pgj_unstuff_masks16 mask range");
  }
  const int64_t clen = (int64_t)compact.size();
  compact.resize((size_t)clen + STREAM_PAD, 0xFF);
  // word loads need 4-byte alignment
  std::vector<uint32_t> aligned((compact.size() + 3) / 4 + 1);
  std::memcpy(aligned.data(), compact.data(), compact.size());
  PgjStream sv{reinterpret_cast<const uint8_t*>(aligned.data()), clen * 8, rst.data(), (int32_t)rst.size()};
  const int chunk_bits = chunk_bytes * 8;
  const int n_chunks = (int)((sl + chunk_bytes - 1) / chunk_bytes);
  std::vector<PgjChunkState> st[2];
  std::vector<PgjEntry> ent((size_t)n_chunks);
  for (int k = 0; k < 2; ++k) st[k].resize((size_t)n_chunks);
  for (int j = 0; j < n_chunks; ++j)  // D4
    pgj_spec_chunk(sv, im, j, chunk_bits, pgj_overlap_bits(chunk_bits), ent[(size_t)j], st[0][(size_t)j]);
  int used = 0, first_changes = 0, parity = 0;
  for (int r = 1; r <= max_rounds; ++r) {  // D5
    const auto& in = st[(r - 1) & 1];
    auto& ov = st[r & 1];
    int changes = 0;
    for (int j = 0; j < n_chunks; ++j) {
      PgjChunkState mine = in[(size_t)j];
      if (j > 0 && mine.p != -2) {
        const PgjChunkState prev = in[(size_t)j - 1];
        if (prev.p != ent[(size_t)j].p || prev.c != ent[(size_t)j].c) {
          pgj_sync_chunk(sv, im, j, chunk_bits, prev, ent[(size_t)j], mine);
          ++changes;
        }
      }
      ov[(size_t)j] = mine;
    }
    parity = r & 1;
    if (r == 1) first_changes = changes;
    if (changes) used = r;
    else break;
  }
  // D6 + D7
  std::vector<int16_t> coef((size_t)im.total_blocks * 64, 0);
  int anchor = 0, n = 0, d0 = 0, d1 = 0, d2 = 0;
  for (int j = 0; j < n_chunks; ++j) {
    const PgjChunkState& cs = st[parity][(size_t)j];
    const int64_t b0 = (int64_t)j * chunk_bits, b1 = b0 + chunk_bits;
    const int64_t ep = j == 0 ? 0 : st[parity][(size_t)j - 1].p;
    const int ec = j == 0 ? 0 : st[parity][(size_t)j - 1].c;
    if (b0 < sv.n_bits && ep >= 0 && ep < b1) {
      HostBlockSink sink{coef.data(), {0}, 0};
      pgj_span_store(sv, im, ep, ec, b1, std::max(anchor, 0) * im.restart_blocks + n, d0, d1, d2, sink);
    }
    if (cs.p != -2) {
      if (cs.anchor >= 0) { anchor = cs.anchor; n = cs.n; d0 = cs.dc[0]; d1 = cs.dc[1]; d2 = cs.dc[2]; }
      else { n += cs.n; d0 += cs.dc[0]; d1 += cs.dc[1]; d2 += cs.dc[2]; }
    }
  }
  // D8 (+ D9 for colour files)
  if (im.n_comps == 1) {
    for (int by = 0; by < im.comp_bh[0]; ++by)
      for (int bx = 0; bx < im.comp_bw[0]; ++bx) {
        uint8_t px[64];
        pgj_idct_block(coef.data() + ((size_t)by * im.comp_bw[0] + bx) * 64, im.qt[0], px);
        for (int y = 0; y < 8 && by * 8 + y < im.height; ++y)
          for (int x = 0; x < 8 && bx * 8 + x < im.width; ++x) out[(int64_t)(by * 8 + y) * pitch + bx * 8 + x] = px[8 * y + x];
      }
  } else {
    std::vector<uint8_t> plane[3];
    for (int c = 0; c < 3; ++c) {
      const int pw = im.comp_bw[c] * 8;
      plane[c].resize((size_t)pw * im.comp_bh[c] * 8);
      for (int by = 0; by < im.comp_bh[c]; ++by)
        for (int bx = 0; bx < im.comp_bw[c]; ++bx) {
          uint8_t px[64];
          pgj_idct_block(coef.data() + im.comp_coef_off[c] + ((size_t)by * im.comp_bw[c] + bx) * 64, im.qt[c], px);
          for (int y = 0; y < 8; ++y)
            for (int x = 0; x < 8; ++x) plane[c][(size_t)(by * 8 + y) * pw + bx * 8 + x] = px[8 * y + x];
        }
    }
    const int hmax = im.comp_h[0], vmax = im.comp_v[0];
    const int fx = hmax / im.comp_h[1], fy = vmax / im.comp_v[1];
    const int dw = (im.width * im.comp_h[1] + hmax - 1) / hmax, dh = (im.height * im.comp_v[1] + vmax - 1) / vmax;
    for (int y = 0; y < im.height; ++y)
      for (int x = 0; x < im.width; ++x) {
        const int yy = plane[0][(size_t)y * im.comp_bw[0] * 8 + x];
        const int cb = pgj_upsample(plane[1].data(), im.comp_bw[1] * 8, dw, dh, fx, fy, x, y);
        const int cr = pgj_upsample(plane[2].data(), im.comp_bw[2] * 8, dw, dh, fx, fy, x, y);
        uint8_t* o = out + (int64_t)y * pitch + 3 * x;
        pgj_ycc_to_bgr(yy, cb, cr, o[0], o[1], o[2]);
      }
  }
  if (stats) { stats[0] = used; stats[1] = first_changes; stats[2] = n_chunks; stats[3] = (int64_t)rst.size(); }
  return PG_OK;
}
