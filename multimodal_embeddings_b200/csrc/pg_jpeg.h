// pg_jpeg.h — baseline JPEG entropy decoding + inverse DCT as inline host/device code, shared by the CUDA
// kernels (pg_jpeg.cu) and by the host evaluation used in the CPU test-suite (pg_hostcheck_jpeg_decode).
//
// Replaces, for `.jpg` scans, the `cv2.imread(image_path)` that opens every stage of the reference
// (1_doclayout_bboxes.py:381, 2_edge_box_filter.py:195): libjpeg-turbo's sequential Huffman decoder (ITU-T T.81
// Annex F), dequantisation and the "islow" integer IDCT (jidctint.c), restated from the published algorithm.
// Bit-exact against cv2.imdecode (tests).
//
// Parallel entropy decoding.  A Huffman bit stream has no index, so it is cut into fixed CHUNKS of the
// unstuffed stream and decoded in three passes (the self-synchronising scheme of Weissenberger & Schmidt,
// "Massively Parallel Huffman Decoding on GPUs", adapted to JPEG block structure and restart markers):
//   1. every chunk is decoded speculatively, starting a quarter of a chunk BEFORE its first bit as if a block
//      started there; Huffman streams resynchronise within a few code words, so the walk has normally fallen
//      into step by the time it enters the chunk.  A chunk's entry / exit state = (bit position, block-in-MCU
//      index) of the first block starting at or after the chunk's first bit / end;
//   2. sync rounds: where chunk j-1's exit differs from the entry chunk j's exit was computed from, chunk j is
//      decoded again from that exit.  Chunk 0 starts in the true state, so a round that redoes nothing means
//      every state is true.  Restart markers (byte-aligned, DC predictors reset) are known true states and
//      resynchronise a wrong decoder at once;
//   3. a segmented scan over the chunks' block counts and DC-difference sums gives every chunk the absolute
//      block index and DC predictors at its entry; the final pass decodes again and stores coefficients.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define PGJ_HD __host__ __device__ __forceinline__
#else
#define PGJ_HD inline
#endif

constexpr int PGJ_LUT_BITS = 9;
constexpr int PGJ_MAX_BLOCKS_PER_MCU = 10;

struct PgjHuff {
  uint16_t lut[1 << PGJ_LUT_BITS];  // (length << 8) | symbol for codes of <= 9 bits; 0: longer code
  // AC tables, same index (the next 9 bits of the stream):
  uint16_t skip[1 << PGJ_LUT_BITS]; // passes 1-2 need no values: code length + magnitude bits (bits 0-4), run (5-8),
                                    // bit 9 = symbol without magnitude bits (EOB / ZRL); 0: longer code
  int16_t fast[1 << PGJ_LUT_BITS];  // pass 3: code and magnitude bits both inside the 9 bits and |value| < 128:
                                    // (value << 8) | (run << 4) | bits; 0: take the two-step path
  int32_t maxcode[18];              // maxcode[l] = largest code of length l (-1: none); [17] = sentinel
  int32_t valoff[17];               // symbol index of the first code of length l minus that code
  uint8_t vals[256];
};

struct alignas(16) PgjImage {  // copied into shared memory 16 bytes at a time
  int32_t width, height, n_comps;
  int32_t mcus_w, mcus_h, bpm;                  // blocks per MCU
  int32_t blk_comp[PGJ_MAX_BLOCKS_PER_MCU];     // component of block c of an MCU
  int32_t blk_dx[PGJ_MAX_BLOCKS_PER_MCU], blk_dy[PGJ_MAX_BLOCKS_PER_MCU];
  int32_t comp_h[3], comp_v[3];                 // sampling factors
  int32_t comp_bw[3], comp_bh[3];               // blocks per row / column of the component's coefficient plane
  int32_t comp_dc[3], comp_ac[3];               // Huffman table ids (0/1)
  int64_t comp_coef_off[3];                     // int16 elements from the image's coefficient base
  int32_t restart_blocks;                       // blocks per restart interval (0: no restart markers)
  int32_t total_blocks;
  uint16_t qt[3][64];                           // per component, natural order
  PgjHuff huff[2][2];                           // [class: 0 DC, 1 AC][table id]
};

struct PgjStream {
  const uint8_t* bytes;     // unstuffed entropy-coded data, readable 512 bytes beyond n_bits / 8
  int64_t n_bits;
  const int32_t* rst_pos;   // byte positions (in `bytes`) where restart interval k+1 starts, ascending
  int32_t n_rst;
};

struct PgjChunkState {   // 32 bytes
  int64_t p;             // exit: bit position of the first block starting at/after the chunk's end; -1 invalid, -2 inactive
  int32_t c;             // block-in-MCU index there
  int32_t n;             // blocks accepted since the chunk's entry, or since the last restart passed
  int32_t anchor;        // -1: no restart passed inside; k: restart interval k was entered (block index k * restart_blocks)
  int32_t dc[3];         // sums of DC differences per component since entry / the last restart
};

// natural-order index of zigzag position k
#define PGJ_ZIGZAG_TABLE                                                                                             \
  {0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28, \
   35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, \
   62, 63}
#ifdef __CUDACC__
static __device__ const uint8_t pgj_zz_dev[64] = PGJ_ZIGZAG_TABLE;  // global memory: lanes index it divergently
#endif
static const uint8_t pgj_zz_host[64] = PGJ_ZIGZAG_TABLE;
PGJ_HD int pgj_zigzag(int k) {
#ifdef __CUDA_ARCH__
  return __ldg(&pgj_zz_dev[k]);
#else
  return pgj_zz_host[k];
#endif
}

// ---- bit reader over the unstuffed stream (big-endian bit order) ---------------------------------------
#ifdef PG_CHECKED
#include <assert.h>
#define PGJ_ASSERT(c) assert(c)
#else
#define PGJ_ASSERT(c) ((void)0)
#endif
constexpr int PGJ_STREAM_PAD = 1024;  // readable bytes of slack (ones) behind every unstuffed stream

struct PgjBits {
  const uint32_t* w0;  // the stream's first word (4-byte aligned)
  int32_t wi;          // index of the next word to load (a stream is far below 2^31 words).  The refill is the walks'
                       // most frequent divergent block: one address, a load, a byte swap, a shift
  int32_t avail;       // valid bits in buf
  uint64_t buf;        // next bits, left-aligned
#ifdef PG_CHECKED
  int64_t last_word;  // last word inside the stream's slack (checked build only)
  PGJ_HD void bound_bits(int64_t n_bits) { last_word = (n_bits + 8 * (int64_t)PGJ_STREAM_PAD) / 32 - 1; }
#else
  PGJ_HD void bound_bits(int64_t) {}
#endif

  PGJ_HD uint32_t load_be(int32_t w) const {
#ifdef PG_CHECKED
    PGJ_ASSERT(w >= 0 && (int64_t)w <= last_word);
#endif
    const uint32_t v = w0[w];
#ifdef __CUDA_ARCH__
    return __byte_perm(v, 0u, 0x0123);
#else
    return (v >> 24) | ((v >> 8) & 0xFF00u) | ((v << 8) & 0xFF0000u) | (v << 24);
#endif
  }
  PGJ_HD void seek(const uint8_t* b, int64_t p) {
    w0 = reinterpret_cast<const uint32_t*>(b);
    wi = (int32_t)(p >> 5);
    buf = ((uint64_t)load_be(wi) << 32) | (uint64_t)load_be(wi + 1);
    wi += 2;
    const int sh = (int)(p & 31);
    buf <<= sh;
    avail = 64 - sh;
  }
  PGJ_HD void need32() {  // at least 32 valid bits afterwards
    if (avail < 32) {
      buf |= (uint64_t)load_be(wi++) << (32 - avail);
      avail += 32;
    }
  }
  PGJ_HD int64_t pos() const { return (int64_t)wi * 32 - avail; }
  PGJ_HD uint32_t peek16() const { return (uint32_t)(buf >> 48); }
  PGJ_HD void skip(int n) { buf <<= n; avail -= n; }
  PGJ_HD uint32_t take(int n) {  // n in 1..16
    const uint32_t v = (uint32_t)(buf >> (64 - n));
    skip(n);
    return v;
  }
};

// one Huffman symbol; -1: no such code.  Needs >= 16 valid bits.
PGJ_HD int pgj_symbol(PgjBits& br, const PgjHuff& h) {
  const uint32_t look = br.peek16();
  const uint32_t e = h.lut[look >> (16 - PGJ_LUT_BITS)];
  if (e) {
    br.skip((int)(e >> 8));
    return (int)(e & 0xFFu);
  }
  for (int l = PGJ_LUT_BITS + 1; l <= 16; ++l) {
    const int32_t code = (int32_t)(look >> (16 - l));
    if (code <= h.maxcode[l]) {
      br.skip(l);
      return h.vals[(code + h.valoff[l]) & 0xFF];
    }
  }
  return -1;
}

PGJ_HD int pgj_extend(uint32_t v, int s) { return v < (1u << (s - 1)) ? (int)v - (1 << s) + 1 : (int)v; }

// One block, decoded leniently: a decoder that is out of step (pass 1) must keep walking so that it can fall
// into step, so a bit pattern that is no code word costs one bit and a run that overshoots the block ends it.
// Neither happens to a decoder in the true state on a valid stream.  `stop`: the next restart boundary / end of the
// scan — a block that reaches it is not a block (padding, or a decoder out of step) and false is returned.
// VALUES: Emit::dc(diff) / Emit::ac(natural index, value) receive the coefficients (pass 3); without it the AC
// symbols are only stepped over (passes 1-2), one table look-up each.
template <bool VALUES, class Emit>
PGJ_HD bool pgj_block(PgjBits& br, const PgjHuff& dc, const PgjHuff& ac, int& dc_diff, Emit& emit, int64_t stop) {
  int s;
  while (true) {
    br.need32();
    s = pgj_symbol(br, dc);
    if (s >= 0) break;
    br.skip(1);
    if (br.pos() >= stop) return false;
  }
  s &= 15;
  int diff = 0;
  if (s) diff = pgj_extend(br.take(s), s);
  dc_diff = diff;
  emit.dc(diff);
  int k = 1;
  while (k < 64) {  // at most 63 symbols of at most 31 bits: a walk through garbage stays inside the stream's slack
    br.need32();
    const uint32_t look = br.peek16();
    if (!VALUES) {
      const uint32_t e = ac.skip[look >> (16 - PGJ_LUT_BITS)];
      if (e) {
        if (e & 512u) {  // EOB or ZRL
          br.skip((int)(e & 31u));
          if (((e >> 5) & 15u) != 15u) break;
          k += 16;
        } else {
          br.skip((int)(e & 31u));
          k += (int)((e >> 5) & 15u) + 1;  // a run that overshoots ends the block (k >= 64)
        }
        continue;
      }
    } else {
      const int f = ac.fast[look >> (16 - PGJ_LUT_BITS)];
      if (f) {
        k += (f >> 4) & 15;
        br.skip(f & 15);
        if (k > 63) break;
        emit.ac(k, f >> 8);
        ++k;
        continue;
      }
    }
    const int rs = pgj_symbol(br, ac);
    if (rs < 0) {
      br.skip(1);
      if (br.pos() >= stop) return false;
      continue;
    }
    const int r = rs >> 4;
    s = rs & 15;
    if (s == 0) {
      if (r != 15) break;  // EOB
      k += 16;             // ZRL
      continue;
    }
    k += r;
    if (k > 63) { br.skip(s); break; }
    emit.ac(k, pgj_extend(br.take(s), s));
    ++k;
  }
  return true;
}

struct PgjNoEmit {
  PGJ_HD void dc(int) {}
  PGJ_HD void ac(int, int) {}
};

// first restart whose boundary lies beyond bit p
PGJ_HD int pgj_next_restart(const PgjStream& sv, int64_t p) {
  int lo = 0, hi = sv.n_rst;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if ((int64_t)sv.rst_pos[mid] * 8 > p) hi = mid; else lo = mid + 1;
  }
  return lo;
}

// Passes 1 and 2: decode (without storing anything) from state (p, c) until a block starts at or after `limit`.
// Interval ends are found from the bits alone: the block that would cross a restart boundary (or the end of the
// stream) is padding — or the work of a decoder that is out of step — and the decoder moves to the boundary,
// which is a true state.
PGJ_HD void pgj_span(const PgjStream& sv, const PgjImage& im, int64_t p, int c, int64_t limit, int anchor_in,
                     PgjChunkState& out) {
  int n = 0, anchor = anchor_in;
  int d0 = 0, d1 = 0, d2 = 0;
  int r = pgj_next_restart(sv, p - 1);  // first boundary at or beyond p
  if (r < sv.n_rst && (int64_t)sv.rst_pos[r] * 8 == p) {
    // p lies exactly on a boundary: the interval before it ended without padding, or the previous chunk's
    // walk already moved here — either way restart interval r + 1 begins at p (taking it twice changes nothing)
    anchor = r + 1;
    c = 0;
    ++r;
  }
  int64_t bound = r < sv.n_rst ? (int64_t)sv.rst_pos[r] * 8 : sv.n_bits;
  PgjBits br;
  br.bound_bits(sv.n_bits);
  br.seek(sv.bytes, p);
  PgjNoEmit none;
  while (p < limit && p < sv.n_bits) {
    const int comp = im.blk_comp[c];
    int diff = 0;
    const bool ok = pgj_block<false>(br, im.huff[0][im.comp_dc[comp]], im.huff[1][im.comp_ac[comp]], diff, none, bound);
    const int64_t pe = br.pos();
    if (!ok || pe > bound) {
      if (bound >= sv.n_bits) { p = sv.n_bits; c = 0; break; }  // end of the scan
      p = bound; c = 0; anchor = r + 1; n = 0; d0 = d1 = d2 = 0;
      ++r;
      bound = r < sv.n_rst ? (int64_t)sv.rst_pos[r] * 8 : sv.n_bits;
      br.seek(sv.bytes, p);
      continue;
    }
    ++n;
    if (comp == 0) d0 += diff; else if (comp == 1) d1 += diff; else d2 += diff;
    c = c + 1 == im.bpm ? 0 : c + 1;
    p = pe;
  }
  out.p = p; out.c = c; out.n = n; out.anchor = anchor; out.dc[0] = d0; out.dc[1] = d1; out.dc[2] = d2;
}

// The state a chunk was entered in: bit position and block-in-MCU index of the first block starting at or after
// the chunk's first bit.
struct PgjEntry {
  int64_t p;
  int32_t c, pad_;
};

// Pass 1 for chunk j.  The walk starts `overlap_bits` BEFORE the chunk (as if a block started there): by the time
// it crosses the chunk's first bit it has normally fallen into step, so the entry state it reaches there — and
// with it the chunk's exit — is normally the true one, and pass 2 has nothing to redo.
PGJ_HD int pgj_overlap_bits(int chunk_bits) { return chunk_bits / 4 < 256 ? (chunk_bits < 256 ? chunk_bits : 256) : chunk_bits / 4; }

PGJ_HD void pgj_exit_from(const PgjStream& sv, const PgjImage& im, int64_t p, int c, int64_t b1, int anchor_in,
                          PgjChunkState& out) {
  if (p >= b1) {  // a block spans the whole chunk: nothing starts inside it
    out.p = p; out.c = c; out.n = 0; out.anchor = anchor_in; out.dc[0] = out.dc[1] = out.dc[2] = 0;
  } else {
    pgj_span(sv, im, p, c, b1, anchor_in, out);
  }
}

PGJ_HD void pgj_spec_chunk(const PgjStream& sv, const PgjImage& im, int j, int chunk_bits, int overlap_bits, PgjEntry& en,
                           PgjChunkState& out) {
  const int64_t b0 = (int64_t)j * chunk_bits, b1 = b0 + chunk_bits;
  en.p = 0; en.c = 0; en.pad_ = 0;
  if (b0 >= sv.n_bits) {
    out.p = -2; out.c = 0; out.n = 0; out.anchor = -1; out.dc[0] = out.dc[1] = out.dc[2] = 0;
    return;
  }
  if (j > 0) {
    PgjChunkState lead;
    pgj_span(sv, im, b0 - overlap_bits, 0, b0, -1, lead);
    en.p = lead.p; en.c = lead.c;
  }
  pgj_exit_from(sv, im, en.p, en.c, b1, j == 0 ? 0 : -1, out);
}

// Pass 2 for chunk j whose predecessor's exit `prev` differs from the entry its own exit was computed from.
PGJ_HD void pgj_sync_chunk(const PgjStream& sv, const PgjImage& im, int j, int chunk_bits, const PgjChunkState& prev,
                           PgjEntry& en, PgjChunkState& out) {
  en.p = prev.p; en.c = prev.c;
  pgj_exit_from(sv, im, prev.p, prev.c, (int64_t)(j + 1) * chunk_bits, -1, out);
}

// where block `blk` (scan order) of the image keeps its 64 coefficients, in int16 elements from the image's base
PGJ_HD int64_t pgj_block_coef_index(const PgjImage& im, int blk, int c, int& comp) {
  comp = im.blk_comp[c];
  if (im.bpm == 1) return (int64_t)blk * 64;
  const int mcu = blk / im.bpm;
  const int my = mcu / im.mcus_w, mx = mcu - my * im.mcus_w;
  const int bx = mx * im.comp_h[comp] + im.blk_dx[c], by = my * im.comp_v[comp] + im.blk_dy[c];
  return im.comp_coef_off[comp] + ((int64_t)by * im.comp_bw[comp] + bx) * 64;
}

// Pass 3: the same walk from a TRUE state with the absolute block index and the DC predictors known; every block is
// handed to `sink` (begin(pred) / dc(diff) / ac(natural index, value) / end(coefficient index)).  Interval ends are
// taken from the block count here (nothing speculative is left).
template <class Sink>
PGJ_HD void pgj_span_store(const PgjStream& sv, const PgjImage& im, int64_t p, int c, int64_t limit, int blk,
                           int pred0, int pred1, int pred2, Sink& sink) {
  int pred[3] = {pred0, pred1, pred2};
  if (im.restart_blocks && blk > 0 && blk % im.restart_blocks == 0) {
    // the entry lies in the padding behind a completed interval (or already on the boundary): the next block
    // is the first of restart interval k and starts on the k-th boundary with cleared predictors
    const int k = blk / im.restart_blocks;
    if (k - 1 >= sv.n_rst) return;
    p = (int64_t)sv.rst_pos[k - 1] * 8;
    pred[0] = pred[1] = pred[2] = 0;
    c = 0;
  }
  int r = pgj_next_restart(sv, p);
  PgjBits br;
  br.bound_bits(sv.n_bits);
  br.seek(sv.bytes, p);
  while (p < limit && p < sv.n_bits && blk < im.total_blocks) {
    int comp;
    const int64_t ci = pgj_block_coef_index(im, blk, c, comp);
    sink.begin(pred[comp]);
    int diff = 0;
    if (!pgj_block<true>(br, im.huff[0][im.comp_dc[comp]], im.huff[1][im.comp_ac[comp]], diff, sink, sv.n_bits + 64)) return;  // corrupt
    sink.end(ci);
    pred[comp] += diff;
    ++blk;
    c = c + 1 == im.bpm ? 0 : c + 1;
    p = br.pos();
    if (im.restart_blocks && blk % im.restart_blocks == 0) {  // interval complete: skip the padding
      if (r >= sv.n_rst) return;
      p = (int64_t)sv.rst_pos[r] * 8;
      ++r;
      pred[0] = pred[1] = pred[2] = 0;
      c = 0;
      br.seek(sv.bytes, p);
    }
  }
}

// ---- jidctint.c ("islow"): 13-bit constants, PASS1_BITS = 2 --------------------------------------------
PGJ_HD int32_t pgj_descale(int32_t x, int n) { return (x + (1 << (n - 1))) >> n; }

PGJ_HD void pgj_idct_1d(int32_t i0, int32_t i1, int32_t i2, int32_t i3, int32_t i4, int32_t i5, int32_t i6, int32_t i7,
                        int shift, int32_t* o) {
  int32_t z1 = (i2 + i6) * 4433;
  const int32_t tmp2 = z1 + i6 * (-15137);
  const int32_t tmp3 = z1 + i2 * 6270;
  const int32_t tmp0 = (i0 + i4) * 8192, tmp1 = (i0 - i4) * 8192;
  const int32_t tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
  int32_t t0 = i7, t1 = i5, t2 = i3, t3 = i1;
  z1 = t0 + t3;
  int32_t z2 = t1 + t2, z3 = t0 + t2, z4 = t1 + t3;
  const int32_t z5 = (z3 + z4) * 9633;
  t0 *= 2446; t1 *= 16819; t2 *= 25172; t3 *= 12299;
  z1 *= -7373; z2 *= -20995; z3 *= -16069; z4 *= -3196;
  z3 += z5; z4 += z5;
  t0 += z1 + z3; t1 += z2 + z4; t2 += z2 + z3; t3 += z1 + z4;
  o[0] = pgj_descale(tmp10 + t3, shift); o[7] = pgj_descale(tmp10 - t3, shift);
  o[1] = pgj_descale(tmp11 + t2, shift); o[6] = pgj_descale(tmp11 - t2, shift);
  o[2] = pgj_descale(tmp12 + t1, shift); o[5] = pgj_descale(tmp12 - t1, shift);
  o[3] = pgj_descale(tmp13 + t0, shift); o[4] = pgj_descale(tmp13 - t0, shift);
}

PGJ_HD uint8_t pgj_clamp_sample(int32_t x) {
  x += 128;
  return (uint8_t)(x < 0 ? 0 : (x > 255 ? 255 : x));
}

// zigzag position of natural index n (the inverse of the zigzag table).  Blocks are kept in ZIGZAG order between the
// entropy pass and the IDCT: the decoder stores coefficient k where it stands in the stream (no table look-up per
// symbol), and here every index is a compile-time constant once the loops are unrolled.
PGJ_HD constexpr int pgj_nat2zz(int n) {
  constexpr uint8_t t[64] = {0,  1,  5,  6,  14, 15, 27, 28, 2,  4,  7,  13, 16, 26, 29, 42, 3,  8,  12, 17, 25, 30,
                             41, 43, 9,  11, 18, 24, 31, 40, 44, 53, 10, 19, 23, 32, 39, 45, 52, 54, 20, 22, 33, 38,
                             46, 51, 55, 60, 21, 34, 37, 47, 50, 56, 59, 61, 35, 36, 48, 49, 57, 58, 62, 63};
  return t[n];
}

// zz: 64 quantised coefficients in zigzag order; q: quantisation table (natural order); out[8][8] samples
PGJ_HD void pgj_idct_block(const int16_t* zz, const uint16_t* q, uint8_t out[64]) {
  int16_t coef[64];
#pragma unroll
  for (int n = 0; n < 64; ++n) coef[n] = zz[pgj_nat2zz(n)];
  int32_t ws[64];
  {
    // a block with nothing but its DC term (blank paper) is flat: both passes reduce to two rounding shifts
    int any = 0;
#pragma unroll
    for (int k = 1; k < 64; ++k) any |= coef[k];
    if (!any) {
      const int32_t col = pgj_descale((int32_t)coef[0] * q[0] * 8192, 11);  // pass 1, column 0 (every row)
      const uint8_t v = pgj_clamp_sample(pgj_descale(col * 8192, 18));      // pass 2, every sample
#pragma unroll
      for (int k = 0; k < 64; ++k) out[k] = v;
      return;
    }
  }
#pragma unroll
  for (int x = 0; x < 8; ++x) {  // pass 1: columns
    int32_t o[8];
    pgj_idct_1d((int32_t)coef[x] * q[x], (int32_t)coef[8 + x] * q[8 + x], (int32_t)coef[16 + x] * q[16 + x],
                (int32_t)coef[24 + x] * q[24 + x], (int32_t)coef[32 + x] * q[32 + x], (int32_t)coef[40 + x] * q[40 + x],
                (int32_t)coef[48 + x] * q[48 + x], (int32_t)coef[56 + x] * q[56 + x], 11, o);
#pragma unroll
    for (int y = 0; y < 8; ++y) ws[8 * y + x] = o[y];
  }
#pragma unroll
  for (int y = 0; y < 8; ++y) {  // pass 2: rows
    int32_t o[8];
    pgj_idct_1d(ws[8 * y], ws[8 * y + 1], ws[8 * y + 2], ws[8 * y + 3], ws[8 * y + 4], ws[8 * y + 5], ws[8 * y + 6],
                ws[8 * y + 7], 18, o);
#pragma unroll
    for (int x = 0; x < 8; ++x) out[8 * y + x] = pgj_clamp_sample(o[x]);
  }
}

// bytes that leave the entropy-coded segment when it is unstuffed: the zero after a data FF, and both bytes of
// a restart marker FF D0..D7 (prev / next: neighbouring bytes, 0 outside the segment)
PGJ_HD bool pgj_is_rst(uint8_t b) { return b >= 0xD0 && b <= 0xD7; }
PGJ_HD bool pgj_keep_byte(uint8_t prev, uint8_t cur, uint8_t next) {
  if (prev == 0xFF && (cur == 0x00 || pgj_is_rst(cur))) return false;
  if (cur == 0xFF && pgj_is_rst(next)) return false;
  return true;
}
PGJ_HD bool pgj_rst_starts(uint8_t cur, uint8_t next) { return cur == 0xFF && pgj_is_rst(next); }

// The same two tests for sixteen bytes at once (four little-endian words), a byte to a bit: keep / rst bit k belongs to
// byte k.  Bytes [lo, hi) of the sixteen lie inside the segment; byte k + 1 exists inside it while k + 1 < lim
// (lim <= 17); `before` = the segment's byte in front of the sixteen (0 when there is none), `after` = the byte
// behind them.  Byte tests run four to a word: x has a zero byte exactly where the result has 0x80, and a multiply
// gathers the four flags into a nibble.
PGJ_HD uint32_t pgj_zero_bytes(uint32_t x) { return ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x | 0x7F7F7F7Fu); }
PGJ_HD uint32_t pgj_flag_nibble(uint32_t m) { return (m * 0x00204081u) >> 28; }
PGJ_HD void pgj_unstuff_masks16(const uint32_t (&w)[4], uint32_t before, uint32_t after, int lo, int hi, int lim,
                                uint32_t& keep, uint32_t& rst) {
  uint32_t ff = 0u, zr = 0u, rs = 0u;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int q = 0; q < 4; ++q) {
    ff |= pgj_flag_nibble(pgj_zero_bytes(~w[q])) << (4 * q);
    zr |= pgj_flag_nibble(pgj_zero_bytes(w[q])) << (4 * q);
    rs |= pgj_flag_nibble(pgj_zero_bytes((w[q] & 0xF8F8F8F8u) ^ 0xD0D0D0D0u)) << (4 * q);
  }
  const uint32_t in = ((1u << hi) - 1u) & ~((1u << lo) - 1u);
  const uint32_t nx_ok = (1u << (lim > 0 ? lim - 1 : 0)) - 1u;
  const uint32_t rs_next = ((rs >> 1) | ((after & 0xF8u) == 0xD0u ? 0x8000u : 0u)) & nx_ok;
  const uint32_t ff_prev = (((ff & in) << 1) | ((before & 0xFFu) == 0xFFu ? 1u : 0u)) & 0xFFFFu;
  const uint32_t gone = (ff_prev & (zr | rs)) | (ff & rs_next);
  keep = in & ~gone;
  rst = in & ff & rs_next;
}


// ---- colour files: jdsample.c "fancy" (triangle) chroma upsampling + jdcolor.c YCbCr -> RGB ----------------------
// One chroma sample of the full-resolution image at (x, y) from the component's own plane `p` (pitch `pp`), whose
// real extent is dw x dh samples (the rest of the plane is block padding; libjpeg replicates the edge instead).
// fx, fy in {1, 2}: the component's horizontal / vertical subsampling against luma.
PGJ_HD int pgj_upsample(const uint8_t* p, int pp, int dw, int dh, int fx, int fy, int x, int y) {
  if (fx == 1 && fy == 1) return p[(int64_t)y * pp + x];
  if (fy == 1) {  // h2v1: 3/4 nearer + 1/4 farther column, rounding 1 / 2 alternately
    const int cx = x >> 1;
    int nx = (x & 1) ? cx + 1 : cx - 1;
    nx = nx < 0 ? 0 : (nx > dw - 1 ? dw - 1 : nx);
    const uint8_t* row = p + (int64_t)y * pp;
    return (3 * row[cx] + row[nx] + ((x & 1) ? 2 : 1)) >> 2;
  }
  const int cy = y >> 1;
  int ny = (y & 1) ? cy + 1 : cy - 1;
  ny = ny < 0 ? 0 : (ny > dh - 1 ? dh - 1 : ny);
  const uint8_t* r0 = p + (int64_t)cy * pp;
  const uint8_t* r1 = p + (int64_t)ny * pp;
  if (fx == 1) return (3 * r0[x] + r1[x] + ((y & 1) ? 2 : 1)) >> 2;  // h1v2
  const int cx = x >> 1;  // h2v2: column sums 3 * nearer row + farther row, then the same across columns (/16)
  int nx = (x & 1) ? cx + 1 : cx - 1;
  nx = nx < 0 ? 0 : (nx > dw - 1 ? dw - 1 : nx);
  const int a = 3 * r0[cx] + r1[cx], b = 3 * r0[nx] + r1[nx];
  return (3 * a + b + ((x & 1) ? 7 : 8)) >> 4;
}

PGJ_HD uint8_t pgj_clamp8(int v) { return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); }

// jdcolor.c with SCALEBITS = 16: FIX(1.40200) = 91881, FIX(1.77200) = 116130, FIX(0.34414) = 22554, FIX(0.71414) = 46802
PGJ_HD void pgj_ycc_to_bgr(int y, int cb, int cr, uint8_t& b, uint8_t& g, uint8_t& r) {
  cb -= 128; cr -= 128;
  r = pgj_clamp8(y + ((91881 * cr + 32768) >> 16));
  b = pgj_clamp8(y + ((116130 * cb + 32768) >> 16));
  g = pgj_clamp8(y + ((-22554 * cb - 46802 * cr + 32768) >> 16));
}
