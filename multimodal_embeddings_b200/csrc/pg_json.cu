// pg_json.cu — device-side writer of the stage-3 records.
//
// The reference ends stage 3 with `json.dump(result, f, indent=2)` (3_combine_grids.py:441-443) of the
// dict built at 3:282-291: image_path, image_size, parameters, boxes [[x1,y1,x2,y2]...], classes,
// scores, class_names, source_jsons.  At tens of thousands of pages per second that Python call is the
// wall-clock bottleneck of the drop-in CLI, so the documents are laid out here, byte for byte as CPython
// prints them, straight from the merge's kept_idx / n_kept:
//
//   J1 json_format_kernel   one thread per kept box: six float reprs (pg_fmt.h) into 32-byte slots and
//                           the byte length of the box's entry in each of the four arrays
//   J2 json_scan_kernel     one CTA per page: exclusive scans of those lengths -> offsets, section sizes,
//                           document size
//   J3 json_offsets_kernel  exclusive scan of the document sizes -> out_off (documents densely packed)
//   J4 json_emit_kernel     head / separators / tail copied cooperatively, one thread per kept box
//                           writes its four entries
//
// The host supplies the page-invariant text (everything before the first array, everything after the
// last, and the class-name string literals) already JSON-encoded; the numbers never visit the host.
#include <cuda_runtime.h>

#include "pg_common.cuh"
#include "pg_fmt.h"

constexpr int JSON_SLOT = PG_JSON_SLOT_BYTES;  // byte 0: length, bytes 1..24: characters
constexpr int JSON_FMT_THREADS = 128;
constexpr int JSON_SCAN_THREADS = 1024;

struct JsonWs {
  uint8_t* slots;     // [N][6][JSON_SLOT]: x1 y1 x2 y2 class score
  uint32_t* len[4];   // [N] entry length, then (after J2) entry offset inside its section
  int64_t* sec;       // [P][4] section byte totals (entries only)
  int64_t* doc_len;   // [P]
};

__host__ __device__ inline int64_t json_align(int64_t x) { return (x + 255) & ~(int64_t)255; }

JsonWs json_layout(void* ws, int64_t n, int32_t p, int64_t* total) {
  uint8_t* b = static_cast<uint8_t*>(ws);
  int64_t o = 0;
  JsonWs w;
  w.slots = b + o; o += json_align(n * 6 * JSON_SLOT);
  for (int q = 0; q < 4; ++q) { w.len[q] = reinterpret_cast<uint32_t*>(b + o); o += json_align(n * 4); }
  w.sec = reinterpret_cast<int64_t*>(b + o); o += json_align((int64_t)p * 4 * 8);
  w.doc_len = reinterpret_cast<int64_t*>(b + o); o += json_align((int64_t)p * 8);
  if (total) *total = o;
  return w;
}

// the fixed text of json.dump(indent=2) around the four arrays (arrays sit at nesting depth 1)
#define J_SEP_CLASSES "  \"classes\": ["
#define J_SEP_SCORES "  \"scores\": ["
#define J_SEP_NAMES "  \"class_names\": ["
__constant__ char kSepText[3][24] = {J_SEP_CLASSES, J_SEP_SCORES, J_SEP_NAMES};
__constant__ int kSepLen[3] = {sizeof(J_SEP_CLASSES) - 1, sizeof(J_SEP_SCORES) - 1, sizeof(J_SEP_NAMES) - 1};
// a non-empty array closes with "\n  ]," (5 bytes + newline of the next key), an empty one with "],"
__device__ __forceinline__ int close_len(int n) { return n > 0 ? 6 : 3; }  // "\n  ],\n" / "],\n"

__device__ __forceinline__ int page_count(const int32_t* n_kept, const int64_t* page_off, int p) {
  return n_kept ? n_kept[p] : (int)(page_off[p + 1] - page_off[p]);
}

__global__ void __launch_bounds__(JSON_FMT_THREADS) json_format_kernel(
    const double* __restrict__ boxes, const double* __restrict__ classes, const double* __restrict__ scores,
    const int32_t* __restrict__ name_id, const int32_t* __restrict__ kept_idx, const int64_t* __restrict__ page_off,
    const int32_t* __restrict__ n_kept, const int64_t* __restrict__ name_off, JsonWs w) {
  const int p = blockIdx.y;
  const int n = page_count(n_kept, page_off, p);
  const int k = blockIdx.x * JSON_FMT_THREADS + threadIdx.x;
  if (k >= n) return;
  const int64_t pos = page_off[p] + k;
  const int64_t gi = kept_idx ? (int64_t)kept_idx[pos] : pos;
  const uint32_t comma = k < n - 1 ? 1u : 0u;
  // characters go to a shared-memory slot (no local-memory arrays), then out as two 16-byte stores
  __shared__ __align__(16) char fs[JSON_FMT_THREADS][JSON_SLOT];
  char* s = fs[threadIdx.x];
  uint32_t len_box = 0, len_class = 0, len_score = 0;
#pragma unroll 1
  for (int q = 0; q < 6; ++q) {
    const double x = q < 4 ? boxes[4 * gi + q] : (q == 4 ? classes[gi] : scores[gi]);
    const int l = pg_format_double_repr(x, s + 1);
    s[0] = (char)l;
    if (q < 4) len_box += (uint32_t)l; else if (q == 4) len_class = (uint32_t)l; else len_score = (uint32_t)l;
    uint4* dst = reinterpret_cast<uint4*>(w.slots + (pos * 6 + q) * JSON_SLOT);
    dst[0] = reinterpret_cast<const uint4*>(s)[0];
    dst[1] = reinterpret_cast<const uint4*>(s)[1];
  }
  // "\n    [" + 4 x ("\n      " + number) + 3 commas + "\n    ]" [+ ","]
  w.len[0][pos] = 6u + 4u * 7u + 3u + 6u + len_box + comma;
  w.len[1][pos] = 5u + len_class + comma;  // "\n    " + number [+ ","]
  w.len[2][pos] = 5u + len_score + comma;
  const int id = name_id[gi];
  w.len[3][pos] = 5u + (uint32_t)(name_off[id + 1] - name_off[id]) + comma;
}

__global__ void __launch_bounds__(JSON_SCAN_THREADS) json_scan_kernel(
    const int64_t* __restrict__ page_off, const int32_t* __restrict__ n_kept, const int64_t* __restrict__ head_off,
    const int64_t* __restrict__ tail_off, JsonWs w) {
  __shared__ int sm[33];
  const int p = blockIdx.x, tid = threadIdx.x;
  const int n = page_count(n_kept, page_off, p);
  const int64_t base = page_off[p];
  int64_t doc = (head_off[p + 1] - head_off[p]) + (tail_off[p + 1] - tail_off[p]);
  for (int q = 0; q < 4; ++q) {
    int64_t carry = 0;
    for (int c = 0; c < n; c += JSON_SCAN_THREADS) {
      const int k = c + tid;
      const int v = k < n ? (int)w.len[q][base + k] : 0;
      int total;
      const int ex = pg_block_exscan(v, sm, &total);
      if (k < n) w.len[q][base + k] = (uint32_t)(carry + ex);  // a section stays below 4 GB
      carry += total;
    }
    if (tid == 0) w.sec[p * 4 + q] = carry;
    doc += carry + close_len(n) + (q < 3 ? kSepLen[q] : -1);  // the tail brings its own leading newline
  }
  if (tid == 0) w.doc_len[p] = doc;
}

__global__ void __launch_bounds__(1024) json_offsets_kernel(int32_t n_pages, const int64_t* __restrict__ doc_len,
                                                            int64_t* __restrict__ out_off) {
  __shared__ int64_t sm[33];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int64_t carry = 0;
  for (int c = 0; c < n_pages; c += 1024) {
    const int p = c + tid;
    const int64_t v = p < n_pages ? doc_len[p] : 0;
    int64_t inc = v;
    for (int o = 1; o < 32; o <<= 1) {
      const int64_t t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) sm[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      const int64_t ws_ = sm[lane];
      int64_t winc = ws_;
      for (int o = 1; o < 32; o <<= 1) {
        const int64_t t = __shfl_up_sync(0xffffffffu, winc, o);
        if (lane >= o) winc += t;
      }
      sm[lane] = winc - ws_;
      if (lane == 31) sm[32] = winc;
    }
    __syncthreads();
    if (p < n_pages) out_off[p] = carry + sm[warp] + inc - v;
    carry += sm[32];
    __syncthreads();
  }
  if (tid == 0) out_off[n_pages] = carry;
}

__device__ __forceinline__ void copy_bytes(uint8_t* dst, const uint8_t* src, int64_t n, int tid, int nthreads) {
  for (int64_t i = tid; i < n; i += nthreads) dst[i] = src[i];
}
__device__ __forceinline__ uint8_t* put_indent(uint8_t* o, int spaces) {
  *o++ = '\n';
  for (int i = 0; i < spaces; ++i) *o++ = ' ';
  return o;
}
__device__ __forceinline__ uint8_t* put_slot(uint8_t* o, const uint8_t* slot) {
  const int l = slot[0];
  for (int i = 0; i < l; ++i) o[i] = slot[1 + i];
  return o + l;
}

constexpr int JSON_STAGE_BYTES = 20 * 1024;  // 128 boxes x (43 + 4 x 24 + 1) bytes = 17.9 KB at most

__global__ void __launch_bounds__(JSON_FMT_THREADS) json_emit_kernel(
    const int32_t* __restrict__ name_id, const int32_t* __restrict__ kept_idx, const int64_t* __restrict__ page_off,
    const int32_t* __restrict__ n_kept, const uint8_t* __restrict__ text, const int64_t* __restrict__ head_off,
    const int64_t* __restrict__ tail_off, const int64_t* __restrict__ name_off, JsonWs w, uint8_t* __restrict__ out,
    int64_t out_capacity, const int64_t* __restrict__ out_off, int32_t n_pages) {
  __shared__ __align__(16) uint8_t json_stage[JSON_STAGE_BYTES + 16];
  __shared__ __align__(16) uint8_t json_slots[JSON_FMT_THREADS * 6 * JSON_SLOT];
  if (out_off[n_pages] > out_capacity) return;  // reported through out_off[n_pages]; nothing is written
  const int p = blockIdx.y, tid = threadIdx.x;
  const int n = page_count(n_kept, page_off, p);
  const int64_t head_len = head_off[p + 1] - head_off[p], tail_len = tail_off[p + 1] - tail_off[p];
  // section starts inside the document
  int64_t sec_at[4], at = head_len;
  for (int q = 0; q < 4; ++q) {
    sec_at[q] = at;
    at += w.sec[p * 4 + q] + close_len(n) + (q < 3 ? kSepLen[q] : -1);
  }
  uint8_t* doc = out + out_off[p];
  if (blockIdx.x == 0) {  // the page-invariant text
    copy_bytes(doc, text + head_off[p], head_len, tid, JSON_FMT_THREADS);
    for (int q = tid; q < 4; q += JSON_FMT_THREADS) {
      uint8_t* o = doc + sec_at[q] + w.sec[p * 4 + q];
      if (n > 0) { *o++ = '\n'; *o++ = ' '; *o++ = ' '; }
      *o++ = ']';
      *o++ = ',';
      if (q < 3) {
        *o++ = '\n';
        for (int i = 0; i < kSepLen[q]; ++i) *o++ = (uint8_t)kSepText[q][i];
      }
    }
    copy_bytes(doc + at, text + tail_off[p], tail_len, tid, JSON_FMT_THREADS);
  }
  // The entries of this CTA's boxes are contiguous inside each of the four arrays: compose each run in
  // shared memory — at the same 16-byte phase as its destination — and write it out with aligned 16-byte
  // stores (byte stores straight to global cost a 32-byte sector each: 28x the traffic, measured).
  const int k0 = blockIdx.x * JSON_FMT_THREADS;
  if (k0 >= n) return;
  const int cnt = n - k0 < JSON_FMT_THREADS ? n - k0 : JSON_FMT_THREADS;
  const int k = k0 + tid;
  const bool active = tid < cnt;
  const int64_t pos0 = page_off[p] + k0, pos = pos0 + (active ? tid : 0);
  const int64_t gi = kept_idx ? (int64_t)kept_idx[pos] : pos;
  {  // this CTA's number slots: one coalesced 16-byte-vector copy instead of dependent byte loads
    const uint4* src = reinterpret_cast<const uint4*>(w.slots + pos0 * 6 * JSON_SLOT);
    uint4* dst = reinterpret_cast<uint4*>(json_slots);
    for (int i = tid; i < cnt * (6 * JSON_SLOT / 16); i += JSON_FMT_THREADS) dst[i] = src[i];
  }
  __syncthreads();
  const uint8_t* slot = json_slots + tid * 6 * JSON_SLOT;
  const bool comma = k < n - 1;
  for (int q = 0; q < 4; ++q) {
    const int64_t start = w.len[q][pos0];
    const int64_t end = k0 + cnt < n ? (int64_t)w.len[q][pos0 + cnt] : w.sec[p * 4 + q];
    const int64_t bytes = end - start;
    uint8_t* gdst = doc + sec_at[q] + start;
    const uint32_t phase = (uint32_t)(reinterpret_cast<uintptr_t>(gdst) & 15u);
    const bool staged = bytes <= JSON_STAGE_BYTES;  // very long class names: straight to global
    uint8_t* base = staged ? json_stage + phase : gdst;
    if (active) {
      uint8_t* o = put_indent(base + (w.len[q][pos] - start), 4);
      if (q == 0) {  // "[", four coordinates, "]"
        *o++ = '[';
        for (int c = 0; c < 4; ++c) {
          o = put_indent(o, 6);
          o = put_slot(o, slot + c * JSON_SLOT);
          if (c < 3) *o++ = ',';
        }
        o = put_indent(o, 4);
        *o++ = ']';
      } else if (q < 3) {  // class, score
        o = put_slot(o, slot + (3 + q) * JSON_SLOT);
      } else {  // class name (already a JSON string literal)
        const int id = name_id[gi];
        const int64_t a = name_off[id], l = name_off[id + 1] - a;
        for (int64_t i = 0; i < l; ++i) o[i] = text[a + i];
        o += l;
      }
      if (comma) *o++ = ',';
    }
    if (staged) {
      __syncthreads();
      const int64_t lead = bytes < ((16 - phase) & 15) ? bytes : ((16 - phase) & 15);
      const int64_t nvec = (bytes - lead) >> 4;
      const int64_t tail_at = lead + (nvec << 4);
      if (tid < lead) gdst[tid] = base[tid];
      const uint4* sv = reinterpret_cast<const uint4*>(base + lead);
      uint4* gv = reinterpret_cast<uint4*>(gdst + lead);
      for (int64_t i = tid; i < nvec; i += JSON_FMT_THREADS) gv[i] = sv[i];
      if (tid < bytes - tail_at) gdst[tail_at + tid] = base[tail_at + tid];
      __syncthreads();
    }
  }
}

extern "C" int64_t pg_json_workspace_bytes(int64_t n_boxes, int32_t n_pages) {
  int64_t total = 0;
  json_layout(nullptr, n_boxes < 1 ? 1 : n_boxes, n_pages < 1 ? 1 : n_pages, &total);
  return total;
}

extern "C" int pg_json_combined(const double* boxes, const double* classes, const double* scores,
                                const int32_t* name_id, const int32_t* kept_idx, const int64_t* page_off,
                                const int32_t* n_kept, int32_t n_pages, int64_t n_boxes, int32_t max_boxes_per_page,
                                const uint8_t* text, const int64_t* head_off, const int64_t* tail_off,
                                const int64_t* name_off, uint8_t* out, int64_t out_capacity, int64_t* out_off,
                                void* ws, int64_t ws_bytes, void* stream) {
  PG_REQUIRE(n_pages >= 0 && n_boxes >= 0 && max_boxes_per_page >= 0 && out_capacity >= 0, "sizes");
  if (n_pages == 0) return PG_OK;
  PG_REQUIRE(page_off && text && head_off && tail_off && name_off && out && out_off && ws, "null device pointer");
  PG_REQUIRE(n_boxes == 0 || (boxes && classes && scores && name_id), "null device pointer (box arrays)");
  if (ws_bytes < pg_json_workspace_bytes(n_boxes, n_pages)) {
    pg_set_error("workspace: pg_json_combined needs %lld bytes, got %lld",
                 (long long)pg_json_workspace_bytes(n_boxes, n_pages), (long long)ws_bytes);
    return PG_ERR_WORKSPACE;
  }
  cudaStream_t s = (cudaStream_t)stream;
  const JsonWs w = json_layout(ws, n_boxes < 1 ? 1 : n_boxes, n_pages, nullptr);
  const unsigned chunks = (unsigned)((max_boxes_per_page + JSON_FMT_THREADS - 1) / JSON_FMT_THREADS);
  const dim3 grid(chunks < 1 ? 1 : chunks, (unsigned)n_pages);
  if (max_boxes_per_page > 0) {
    json_format_kernel<<<grid, JSON_FMT_THREADS, 0, s>>>(boxes, classes, scores, name_id, kept_idx, page_off, n_kept,
                                                         name_off, w);
    PG_LAUNCH_CHECK();
  }
  json_scan_kernel<<<n_pages, JSON_SCAN_THREADS, 0, s>>>(page_off, n_kept, head_off, tail_off, w);
  PG_LAUNCH_CHECK();
  json_offsets_kernel<<<1, 1024, 0, s>>>(n_pages, w.doc_len, out_off);
  PG_LAUNCH_CHECK();
  json_emit_kernel<<<grid, JSON_FMT_THREADS, 0, s>>>(name_id, kept_idx, page_off, n_kept, text, head_off, tail_off,
                                                     name_off, w, out, out_capacity, out_off, n_pages);
  PG_LAUNCH_CHECK();
  return PG_OK;
}

extern "C" int32_t pg_hostcheck_format_double(double x, char* buf) { return pg_format_double_repr(x, buf); }

extern "C" int64_t pg_hostcheck_format_doubles(const double* x, int64_t n, char* buf /*n * 24*/, int32_t* len) {
  int64_t total = 0;
  for (int64_t i = 0; i < n; ++i) {
    len[i] = pg_format_double_repr(x[i], buf + i * PG_FMT_MAX_DOUBLE);
    total += len[i];
  }
  return total;
}

extern "C" int64_t pg_hostcheck_parse_numbers(const char* text, int64_t text_len, const int64_t* tok_off, int64_t n,
                                              double* out, int32_t* consumed) {
  int64_t ok = 0;
  for (int64_t i = 0; i < n; ++i) {
    const int64_t avail = text_len - tok_off[i];
    out[i] = 0.0;
    consumed[i] = avail > 0 ? pg_parse_json_number(text + tok_off[i], (int)(avail > 64 ? 64 : avail), &out[i]) : 0;
    ok += consumed[i] > 0;
  }
  return ok;
}

// =============================================================================================
// Reader half: the numbers of a record's "boxes" / "classes" / "scores" arrays, text -> f64
// on the device.  Replaces the float() calls inside json.load of the reference's stage-4/5 readers
// (4_extract_median_widths.py:103-151, 5_detect_column_centers.py:337-400) for the arrays that make up
// > 95 % of a stage-3 record.  The caller locates the arrays (three byte ranges per record; they hold
// nothing but numbers, brackets, commas and whitespace) and parses the few strings on the host.
//   R1 json_tok_count_kernel  one thread per byte: is it the first byte of a number token?  count per 256 B
//   R2 json_tok_scan_kernel   exclusive scan of the block counts -> where each block's values go; per-range offsets
//   R3 json_tok_parse_kernel  the flagged threads convert their token (pg_fmt.h) and store it in text order
// =============================================================================================
constexpr int JSON_TOK_THREADS = 256;
constexpr int JSON_TOK_BPT = PG_JSON_PARSE_BLOCK / JSON_TOK_THREADS;  // consecutive bytes per thread (8)

__device__ __forceinline__ bool json_is_token_start(uint8_t c, uint8_t pv, bool first) {
  if (!((c >= '0' && c <= '9') || c == '-' || c == 'N' || c == 'I')) return false;
  return first || pv == ' ' || pv == '\n' || pv == '[' || pv == ',' || pv == '\t' || pv == '\r';
}

// the range that owns global block b: the last r with blk_off[r] <= b
__device__ __forceinline__ int json_range_of_block(const int64_t* __restrict__ blk_off, int n_ranges, int64_t b) {
  int lo = 0, hi = n_ranges - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (blk_off[mid] <= b) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// bit k of the result: byte i0 + k starts a number token
__device__ __forceinline__ uint32_t json_thread_flags(const uint8_t* __restrict__ text, int64_t i0, int64_t begin,
                                                      int64_t end) {
  uint32_t bits = 0;
  if (i0 >= end) return 0;
  uint8_t pv = i0 > begin ? text[i0 - 1] : (uint8_t)' ';
#pragma unroll
  for (int k = 0; k < JSON_TOK_BPT; ++k) {
    const int64_t i = i0 + k;
    if (i < end) {
      const uint8_t c = text[i];
      if (json_is_token_start(c, pv, i == begin)) bits |= 1u << k;
      pv = c;
    }
  }
  return bits;
}

__global__ void __launch_bounds__(JSON_TOK_THREADS) json_tok_count_kernel(
    const uint8_t* __restrict__ text, const int64_t* __restrict__ ranges, const int64_t* __restrict__ blk_off,
    int n_ranges, int32_t* __restrict__ block_count) {
  __shared__ int sm[33];
  const int64_t b = blockIdx.x;
  const int r = json_range_of_block(blk_off, n_ranges, b);
  const int64_t begin = ranges[2 * r], end = ranges[2 * r + 1];
  const int64_t i0 = begin + (b - blk_off[r]) * PG_JSON_PARSE_BLOCK + (int64_t)threadIdx.x * JSON_TOK_BPT;
  const int cnt = __popc(json_thread_flags(text, i0, begin, end));
  int total;
  pg_block_exscan(cnt, sm, &total);
  if (threadIdx.x == 0) block_count[b] = total;
}

__global__ void __launch_bounds__(1024) json_tok_scan_kernel(int64_t n_blocks, const int32_t* __restrict__ block_count,
                                                             int64_t* __restrict__ block_base,
                                                             const int64_t* __restrict__ blk_off, int n_ranges,
                                                             int64_t* __restrict__ val_off) {
  __shared__ int sm[33];
  __shared__ int64_t carry_s;
  const int tid = threadIdx.x;
  int64_t carry = 0;
  for (int64_t c = 0; c < n_blocks; c += 1024) {
    const int64_t b = c + tid;
    const int v = b < n_blocks ? block_count[b] : 0;
    int total;
    const int ex = pg_block_exscan(v, sm, &total);
    if (b < n_blocks) block_base[b] = carry + ex;
    carry += total;
  }
  if (tid == 0) carry_s = carry;
  __syncthreads();
  for (int r = tid; r <= n_ranges; r += 1024)
    val_off[r] = (r == n_ranges || blk_off[r] >= n_blocks) ? carry_s : block_base[blk_off[r]];
}

__global__ void __launch_bounds__(JSON_TOK_THREADS) json_tok_parse_kernel(
    const uint8_t* __restrict__ text, const int64_t* __restrict__ ranges, const int64_t* __restrict__ blk_off,
    int n_ranges, const int64_t* __restrict__ block_base, double* __restrict__ values, int64_t capacity,
    int32_t* __restrict__ n_bad, int32_t flags) {
  __shared__ int sm[33];
  const int64_t b = blockIdx.x;
  const int r = json_range_of_block(blk_off, n_ranges, b);
  const int64_t begin = ranges[2 * r], end = ranges[2 * r + 1];
  const int64_t i0 = begin + (b - blk_off[r]) * PG_JSON_PARSE_BLOCK + (int64_t)threadIdx.x * JSON_TOK_BPT;
  uint32_t bits = json_thread_flags(text, i0, begin, end);
  int total;
  const int ex = pg_block_exscan(__popc(bits), sm, &total);
  int64_t at = block_base[b] + ex;
  while (bits) {
    const int k = __ffs(bits) - 1;
    bits &= bits - 1;
    const int64_t i = i0 + k;
    double v = 0.0;
    const int64_t avail = end - i;
    bool is_int;
    const int used = pg_parse_json_number(reinterpret_cast<const char*>(text + i), (int)(avail > 48 ? 48 : avail), &v, &is_int);
    if (used == 0 || (is_int && (flags & PG_JSON_INT_LITERALS_TO_HOST))) atomicAdd(&n_bad[r], 1);
    if (at < capacity) values[at] = v;
    ++at;
  }
}

extern "C" int32_t pg_json_parse_block_bytes(void) { return PG_JSON_PARSE_BLOCK; }
extern "C" int64_t pg_json_parse_workspace_bytes(int64_t total_blocks) {
  const int64_t nb = total_blocks < 1 ? 1 : total_blocks;
  return json_align(nb * 4) + json_align(nb * 8);
}

extern "C" int pg_json_parse_numbers(const uint8_t* text, const int64_t* ranges, int32_t n_ranges,
                                     const int64_t* range_block_off, int64_t total_blocks, double* values,
                                     int64_t capacity, int64_t* val_off, int32_t* n_bad, int32_t flags, void* ws,
                                     int64_t ws_bytes, void* stream) {
  PG_REQUIRE(n_ranges >= 0 && total_blocks >= 0 && capacity >= 0, "sizes");
  if (n_ranges == 0) return PG_OK;
  PG_REQUIRE(text && ranges && range_block_off && val_off && n_bad && ws && (values || capacity == 0), "null device pointer");
  PG_REQUIRE(total_blocks <= 0x7fffffffll, "more than 2^31 blocks of text in one call");
  if (ws_bytes < pg_json_parse_workspace_bytes(total_blocks)) {
    pg_set_error("workspace: pg_json_parse_numbers needs %lld bytes, got %lld",
                 (long long)pg_json_parse_workspace_bytes(total_blocks), (long long)ws_bytes);
    return PG_ERR_WORKSPACE;
  }
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t nb = total_blocks < 1 ? 1 : total_blocks;
  int32_t* block_count = static_cast<int32_t*>(ws);
  int64_t* block_base = reinterpret_cast<int64_t*>(static_cast<uint8_t*>(ws) + json_align(nb * 4));
  PG_CUDA_TRY(cudaMemsetAsync(n_bad, 0, sizeof(int32_t) * n_ranges, s));
  if (total_blocks > 0) {
    json_tok_count_kernel<<<(unsigned)total_blocks, JSON_TOK_THREADS, 0, s>>>(text, ranges, range_block_off, n_ranges,
                                                                              block_count);
    PG_LAUNCH_CHECK();
  }
  json_tok_scan_kernel<<<1, 1024, 0, s>>>(total_blocks, block_count, block_base, range_block_off, n_ranges, val_off);
  PG_LAUNCH_CHECK();
  if (total_blocks > 0) {
    json_tok_parse_kernel<<<(unsigned)total_blocks, JSON_TOK_THREADS, 0, s>>>(text, ranges, range_block_off, n_ranges,
                                                                              block_base, values, capacity, n_bad, flags);
    PG_LAUNCH_CHECK();
  }
  return PG_OK;
}

// =============================================================================================
// Generic form of the writer: documents as a flat list of SEGMENTS — a piece of host-encoded text followed
// by one array printed from device data — so that any of the reference's json.dump(indent=2) schemas
// (the nested cells[].regions of stages 1/2 included) can be laid out on the device.  A segment prints
//     head text, then  "[]"                                   if it has no elements,
//                      "[" entries "\n" <indent-2 spaces> "]"  otherwise,
// where an entry is, at `indent` spaces:  BOX4   "\n<indent>[" 4 x "\n<indent+2>number"(",") "\n<indent>]"
//                                          SCALAR "\n<indent>number"      NAME "\n<indent>"string literal"
// each followed by "," unless it is the last.  KIND_TEXT segments are text only (the end of a file).
// Numbers are formatted twice (once for their length, once in place) instead of going through slots: this
// path serves the command-line stages, whose volume is a few MB per page.
//   S1 seg_len_kernel    one thread per element: its entry length
//   S2 seg_scan_kernel   one CTA per segment: entry offsets, segment size
//   S3 json_offsets_kernel (shared with J3): segment offsets in the output
//   S4 seg_head_kernel   one CTA per segment: head text and brackets;  seg_emit_kernel: one thread per element
// =============================================================================================
constexpr int SEG_MAX_DATA = 8;
struct SegData {
  const void* ptr[SEG_MAX_DATA];
};

__device__ __forceinline__ int seg_of_elem(const int64_t* __restrict__ elem_off, int n_segs, int64_t e) {
  int lo = 0, hi = n_segs - 1;
  while (lo < hi) {  // the last segment whose first element is <= e (empty segments share an offset with a later one)
    const int mid = (lo + hi + 1) >> 1;
    if (elem_off[mid] <= e) lo = mid; else hi = mid - 1;
  }
  return lo;
}

__device__ __forceinline__ int64_t seg_body_len(const PgJsonSegment& sg, int64_t entries) {
  if (sg.kind == PG_JSON_KIND_TEXT) return 0;
  return sg.count == 0 ? 2 : 1 + entries + 1 + (sg.indent - 2) + 1;
}

__global__ void __launch_bounds__(128) seg_len_kernel(const PgJsonSegment* __restrict__ segs, int n_segs,
                                                      const int64_t* __restrict__ elem_off, int64_t n_elems, SegData data,
                                                      const int32_t* __restrict__ kept_idx,
                                                      const int64_t* __restrict__ name_off, uint32_t* __restrict__ len) {
  __shared__ char scratch[128][PG_FMT_MAX_DOUBLE + 8];
  const int64_t e = (int64_t)blockIdx.x * 128 + threadIdx.x;
  if (e >= n_elems) return;
  const int s = seg_of_elem(elem_off, n_segs, e);
  const PgJsonSegment sg = segs[s];
  const int64_t k = e - elem_off[s];
  const int64_t pos = sg.start + k;
  const int64_t gi = kept_idx ? (int64_t)kept_idx[pos] : pos;
  const uint32_t comma = k < sg.count - 1 ? 1u : 0u;
  const uint32_t ind = (uint32_t)sg.indent;
  char* buf = scratch[threadIdx.x];
  uint32_t l;
  if (sg.kind == PG_JSON_KIND_BOX4) {
    const double* b = static_cast<const double*>(data.ptr[sg.data_id]) + 4 * gi;
    l = (2u + ind) + 4u * (3u + ind) + 3u + (2u + ind);
    for (int c = 0; c < 4; ++c) l += (uint32_t)pg_format_double_repr(b[c], buf);
  } else if (sg.kind == PG_JSON_KIND_SCALAR) {
    l = 1u + ind + (uint32_t)pg_format_double_repr(static_cast<const double*>(data.ptr[sg.data_id])[gi], buf);
  } else {
    const int id = static_cast<const int32_t*>(data.ptr[sg.data_id])[gi];
    l = 1u + ind + (uint32_t)(name_off[id + 1] - name_off[id]);
  }
  len[e] = l + comma;
}

__global__ void __launch_bounds__(256) seg_scan_kernel(const PgJsonSegment* __restrict__ segs,
                                                       const int64_t* __restrict__ elem_off, uint32_t* __restrict__ len,
                                                       int64_t* __restrict__ seg_len) {
  __shared__ int sm[33];
  const int s = blockIdx.x, tid = threadIdx.x;
  const PgJsonSegment sg = segs[s];
  const int64_t base = elem_off[s];
  const int n = sg.kind == PG_JSON_KIND_TEXT ? 0 : sg.count;
  int64_t carry = 0;
  for (int c = 0; c < n; c += 256) {
    const int k = c + tid;
    const int v = k < n ? (int)len[base + k] : 0;
    int total;
    const int ex = pg_block_exscan(v, sm, &total);
    if (k < n) len[base + k] = (uint32_t)(carry + ex);
    carry += total;
  }
  if (tid == 0) seg_len[s] = (sg.head_end - sg.head_begin) + seg_body_len(sg, carry);
}

__global__ void __launch_bounds__(128) seg_head_kernel(const PgJsonSegment* __restrict__ segs, int n_segs,
                                                       const uint8_t* __restrict__ text, uint8_t* __restrict__ out,
                                                       int64_t capacity, const int64_t* __restrict__ seg_out_off) {
  if (seg_out_off[n_segs] > capacity) return;
  const int s = blockIdx.x, tid = threadIdx.x;
  const PgJsonSegment sg = segs[s];
  uint8_t* o = out + seg_out_off[s];
  const int64_t hl = sg.head_end - sg.head_begin;
  for (int64_t i = tid; i < hl; i += 128) o[i] = text[sg.head_begin + i];
  if (tid == 0 && sg.kind != PG_JSON_KIND_TEXT) {
    o[hl] = '[';
    uint8_t* end = out + seg_out_off[s + 1];
    end[-1] = ']';
    if (sg.count > 0) {
      uint8_t* q = end - 1 - (sg.indent - 2) - 1;
      *q++ = '\n';
      for (int i = 0; i < sg.indent - 2; ++i) *q++ = ' ';
    }
  }
}

__global__ void __launch_bounds__(128) seg_emit_kernel(const PgJsonSegment* __restrict__ segs, int n_segs,
                                                       const int64_t* __restrict__ elem_off, int64_t n_elems, SegData data,
                                                       const int32_t* __restrict__ kept_idx, const uint8_t* __restrict__ text,
                                                       const int64_t* __restrict__ name_off,
                                                       const uint32_t* __restrict__ len, uint8_t* __restrict__ out,
                                                       int64_t capacity, const int64_t* __restrict__ seg_out_off) {
  if (seg_out_off[n_segs] > capacity) return;
  const int64_t e = (int64_t)blockIdx.x * 128 + threadIdx.x;
  if (e >= n_elems) return;
  const int s = seg_of_elem(elem_off, n_segs, e);
  const PgJsonSegment sg = segs[s];
  const int64_t k = e - elem_off[s];
  const int64_t pos = sg.start + k;
  const int64_t gi = kept_idx ? (int64_t)kept_idx[pos] : pos;
  uint8_t* o = out + seg_out_off[s] + (sg.head_end - sg.head_begin) + 1 + len[e];
  o = put_indent(o, sg.indent);
  if (sg.kind == PG_JSON_KIND_BOX4) {
    const double* b = static_cast<const double*>(data.ptr[sg.data_id]) + 4 * gi;
    *o++ = '[';
    for (int c = 0; c < 4; ++c) {
      o = put_indent(o, sg.indent + 2);
      o += pg_format_double_repr(b[c], reinterpret_cast<char*>(o));
      if (c < 3) *o++ = ',';
    }
    o = put_indent(o, sg.indent);
    *o++ = ']';
  } else if (sg.kind == PG_JSON_KIND_SCALAR) {
    o += pg_format_double_repr(static_cast<const double*>(data.ptr[sg.data_id])[gi], reinterpret_cast<char*>(o));
  } else {
    const int id = static_cast<const int32_t*>(data.ptr[sg.data_id])[gi];
    const int64_t a = name_off[id], l = name_off[id + 1] - a;
    for (int64_t i = 0; i < l; ++i) o[i] = text[a + i];
    o += l;
  }
  if (k < sg.count - 1) *o++ = ',';
}

extern "C" int64_t pg_json_segments_workspace_bytes(int64_t n_elems, int32_t n_segs) {
  return json_align((n_elems < 1 ? 1 : n_elems) * 4) + json_align((int64_t)(n_segs < 1 ? 1 : n_segs) * 8);
}

extern "C" int pg_json_segments(const PgJsonSegment* segs, int32_t n_segs, const int64_t* elem_off, int64_t n_elems,
                                const void* const* data, int32_t n_data, const int32_t* kept_idx, const uint8_t* text,
                                const int64_t* name_off, uint8_t* out, int64_t out_capacity, int64_t* seg_out_off,
                                void* ws, int64_t ws_bytes, void* stream) {
  PG_REQUIRE(n_segs >= 0 && n_elems >= 0 && n_data >= 0 && n_data <= SEG_MAX_DATA && out_capacity >= 0, "sizes");
  if (n_segs == 0) return PG_OK;
  PG_REQUIRE(segs && elem_off && text && out && seg_out_off && ws && (data || n_data == 0), "null pointer");
  if (ws_bytes < pg_json_segments_workspace_bytes(n_elems, n_segs)) {
    pg_set_error("workspace: pg_json_segments needs %lld bytes, got %lld",
                 (long long)pg_json_segments_workspace_bytes(n_elems, n_segs), (long long)ws_bytes);
    return PG_ERR_WORKSPACE;
  }
  cudaStream_t s = (cudaStream_t)stream;
  SegData d;
  for (int i = 0; i < SEG_MAX_DATA; ++i) d.ptr[i] = i < n_data ? data[i] : nullptr;
  uint32_t* len = static_cast<uint32_t*>(ws);
  int64_t* seg_len = reinterpret_cast<int64_t*>(static_cast<uint8_t*>(ws) + json_align((n_elems < 1 ? 1 : n_elems) * 4));
  const unsigned eg = (unsigned)((n_elems + 127) / 128);
  if (n_elems > 0) {
    seg_len_kernel<<<eg, 128, 0, s>>>(segs, n_segs, elem_off, n_elems, d, kept_idx, name_off, len);
    PG_LAUNCH_CHECK();
  }
  seg_scan_kernel<<<n_segs, 256, 0, s>>>(segs, elem_off, len, seg_len);
  PG_LAUNCH_CHECK();
  json_offsets_kernel<<<1, 1024, 0, s>>>(n_segs, seg_len, seg_out_off);
  PG_LAUNCH_CHECK();
  seg_head_kernel<<<n_segs, 128, 0, s>>>(segs, n_segs, text, out, out_capacity, seg_out_off);
  PG_LAUNCH_CHECK();
  if (n_elems > 0) {
    seg_emit_kernel<<<eg, 128, 0, s>>>(segs, n_segs, elem_off, n_elems, d, kept_idx, text, name_off, len, out,
                                       out_capacity, seg_out_off);
    PG_LAUNCH_CHECK();
  }
  return PG_OK;
}
