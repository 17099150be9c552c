// pg_tiler.cu — K1: overlapped grid tiling + letterbox resize + normalise, uint8 BGR page ->
// planar RGB fp16 tiles.
//
// Reference path replaced: split_image_into_grid (1_doclayout_bboxes.py:366-444) feeding
// YOLODocumentLayoutDetector.detect_regions (1_doclayout_bboxes.py:191-210), i.e. the third-party
// LetterBox -> cv2.resize(INTER_LINEAR) -> copyMakeBorder(114) -> BGR->RGB -> CHW -> /255 chain.
//
// Design (B200): HBM-bound streaming kernel, no tensor cores.  CTAs of 9 warps with a bounded
// lifetime: warp 8 is a producer whose elected lane claims work items (bands of 16 output rows) from
// a shared counter and stages, per output row, the two source rows it needs with TMA bulk copies
// (cp.async.bulk -> UBLKCP) into a 3-deep shared-memory ring guarded by full/empty mbarriers, plus a
// 16-byte message (chunk, page, row, vertical coefficients); warps 0-7 consume: 3x LDS.32 per source
// row and pixel, PRMT + IDP.2A for the 11-bit horizontal pass, integer vertical pass (bit-exact cv2
// model, pg_math.h) finished two pixels to a register, a half2 FMA for the /255 and streaming half2
// stores into the three colour planes.  Source rows that no output row samples (scale > 2) are never
// read.  Tiles with long source rows are cut into column chunks so that every grid keeps 4 CTAs per
// SM, and the work items of ALL grids of a page are ordered by the source row they start at and
// claimed dynamically: everything that samples the same part of the page (adjacent tiles, the 20 %
// overlaps, the other grids) is in flight together and shares it through L2.  A CTA retires after
// <= 4 items, which lets the box-stage kernels on a higher-priority stream share the SMs.  Measured
// on B200: 1.05 of the copy-measured HBM peak (algorithmic bytes) for 64 pages of 8000x6000 at 4x4,
// DRAM traffic 0.94x the algorithmic bytes at 6.5 TB/s; the reference's default 30-tile grid set
// reads every page byte from DRAM exactly once.
#include <cuda_fp16.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "pg_common.cuh"

// ------------------------------------------------------------------------------------------
// error plumbing (shared by all translation units)
static thread_local char g_err[512] = "";
void pg_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
extern "C" const char* pg_last_error(void) { return g_err; }
extern "C" int pg_version(void) { return 100; }
extern "C" int pg_build_checked(void) {
#ifdef PG_CHECKED
  return 1;
#else
  return 0;
#endif
}
extern "C" int pg_device_info(int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor) {
  int dev = 0;
  PG_CUDA_TRY(cudaGetDevice(&dev));
  cudaDeviceProp p;
  PG_CUDA_TRY(cudaGetDeviceProperties(&p, dev));
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  return PG_OK;
}

// ------------------------------------------------------------------------------------------
// plan
constexpr int TL_BAND = 16;             // output rows per work item
constexpr uint32_t TL_PADMARK = 0xFFFFFFFFu;
constexpr int TL_PAD_VALUE = 114;

// A tile as the kernels see it.  The pipeline kernel works on COLUMN CHUNKS of a tile (chunk_x0, chunk_w):
// a tile whose source rows are too long for four resident CTAs per SM (TL_MAX_ROW_BYTES) is cut into
// chunks of output columns, each staging only the source span it samples (x0 = first source pixel of the
// chunk, x-table rebased to it).  A full tile is the chunk [0, out_w) — what the direct kernel uses.
struct TileDev {
  int32_t x0, y0, src_w, src_h;
  int32_t new_w, new_h, pad_l, pad_t, out_w, out_h;
  int32_t xtab_off, ytab_off;
  int32_t row_bytes;  // bulk-copy size per source row (multiple of 16)
  int32_t row_skew;   // byte offset of source pixel 0 inside the staged row
  int32_t chunk_x0, chunk_w;  // output columns [chunk_x0, chunk_x0 + chunk_w) of the tile (even / multiple of 64)
  int32_t has_xpad, pad_;     // the chunk holds 114-valued pad columns
  int64_t out_off;    // fp16 elements from the page's output base
};
constexpr int TL_MAX_ROW_BYTES = 9216;
// Work counters are per launch: launch k of a plan (or batch) uses slot k % 8 of a small ring, which its last CTA to
// retire re-arms, so launches of one plan that overlap on different streams never share a counter (up to 8 in
// flight).  A launch recorded into a CUDA graph keeps its slot for every replay; those take slots 8..15 so that a
// replaying graph and eager launches of the same plan cannot meet either.
constexpr int TL_COUNTER_SLOTS = 16;
static unsigned long long* counter_slot(unsigned long long* base, std::atomic<uint32_t>& seq, cudaStream_t s) {
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  const bool capturing = cudaStreamIsCapturing(s, &st) == cudaSuccess && st == cudaStreamCaptureStatusActive;
  const uint32_t k = seq.fetch_add(1u) % (TL_COUNTER_SLOTS / 2);
  return base + 2 * (k + (capturing ? TL_COUNTER_SLOTS / 2 : 0));
}  // 3 stages x 2 rows x (9216+128) B x 4 CTAs fits the 227 KB of an SM

struct PgTilePlan {
  int32_t page_w = 0, page_h = 0, imgsz = 0;
  int32_t channels = 3;  // 3: BGR interleaved pages (cv2.imread); 1: one grey plane (a greyscale scan decoded on the device)
  std::vector<PgTileInfo> info;
  std::vector<TileDev> tiles;       // column chunks (pipeline kernel); items index this table
  std::vector<TileDev> tiles_full;  // whole tiles (direct validation kernel), one per PgTileInfo
  std::vector<uint2> xtab;       // per chunk, chunk_w entries: {3*(s0 - chunk source start), a0 | a1<<16} or {0, PADMARK}
  std::vector<uint2> xtab_full;  // per tile, out_w entries relative to the tile's x0
  std::vector<int4> ytab;   // per tile, new_h entries: {s0, s1, b0, b1}
  std::vector<int4> items;  // {chunk, oy0, nrows, 0}, all grids together, ordered by the source row they start at
  int64_t out_elems = 0;
  int32_t max_row_bytes = 0;  // widest staged row over the chunks
  int32_t max_out_w = 0;      // widest chunk (px)
  // device mirrors (lazy)
  int device = -1;
  TileDev* d_tiles = nullptr;
  TileDev* d_tiles_full = nullptr;
  uint2* d_xtab = nullptr;
  uint2* d_xtab_full = nullptr;
  int4* d_ytab = nullptr;
  int4* d_items = nullptr;
  unsigned long long* d_counters = nullptr;  // TL_COUNTER_SLOTS x {work-item counter, retired-CTA counter}
  std::atomic<uint32_t> launch_seq{0};
};

static double py_round_half_even(double v) { return std::nearbyint(v); }

extern "C" int pg_tile_plan_create_ex(int32_t page_w, int32_t page_h, int32_t channels, const int32_t* grid_rows,
                                      const int32_t* grid_cols, int32_t n_grids, double overlap_percentage,
                                      int32_t imgsz, int32_t stride, int32_t auto_pad, int32_t scaleup,
                                      PgTilePlan** plan_out);
extern "C" int pg_tile_plan_create(int32_t page_w, int32_t page_h, const int32_t* grid_rows,
                                   const int32_t* grid_cols, int32_t n_grids, double overlap_percentage,
                                   int32_t imgsz, int32_t stride, int32_t auto_pad, int32_t scaleup,
                                   PgTilePlan** plan_out) {
  return pg_tile_plan_create_ex(page_w, page_h, 3, grid_rows, grid_cols, n_grids, overlap_percentage, imgsz, stride,
                                auto_pad, scaleup, plan_out);
}

extern "C" int pg_tile_plan_create_ex(int32_t page_w, int32_t page_h, int32_t channels, const int32_t* grid_rows,
                                      const int32_t* grid_cols, int32_t n_grids, double overlap_percentage,
                                      int32_t imgsz, int32_t stride, int32_t auto_pad, int32_t scaleup,
                                      PgTilePlan** plan_out) {
  PG_REQUIRE(plan_out != nullptr, "plan");
  PG_REQUIRE(channels == 1 || channels == 3, "channels must be 1 (grey plane) or 3 (BGR)");
  const int CHN = channels;
  PG_REQUIRE(page_w > 0 && page_h > 0, "page size");
  PG_REQUIRE(n_grids > 0 && grid_rows && grid_cols, "grids");
  PG_REQUIRE(imgsz > 0 && stride > 0 && imgsz % 2 == 0, "imgsz/stride");
  auto* plan = new PgTilePlan();
  plan->page_w = page_w;
  plan->page_h = page_h;
  plan->imgsz = imgsz;
  plan->channels = channels;
  int64_t out_off = 0;
  int max_row_bytes = TL_MAX_ROW_BYTES;
  if (const char* e = getenv("PG_TILER_MAX_ROW_BYTES")) max_row_bytes = std::max(64, atoi(e));  // test / tuning knob
  for (int g = 0; g < n_grids; ++g) {
    const int rows = grid_rows[g], cols = grid_cols[g];
    if (rows <= 0 || cols <= 0) {
      delete plan;
      pg_set_error("invalid argument: grid %d is %dx%d", g, rows, cols);
      return PG_ERR_INVALID;
    }
    // 1_doclayout_bboxes.py:388-394 (Python doubles)
    const double bw = (double)page_w / (double)cols;
    const double bh = (double)page_h / (double)rows;
    const double ox = bw * (overlap_percentage / 100.0);
    const double oy = bh * (overlap_percentage / 100.0);
    for (int r = 0; r < rows; ++r) {
      for (int c = 0; c < cols; ++c) {
        double xs = c * bw;                       // :401-403
        if (c > 0) xs -= ox;
        double ys = r * bh;                       // :405-407
        if (r > 0) ys -= oy;
        double xe = (c + 1) * bw;                 // :409-411
        if (c < cols - 1) xe += ox;
        double ye = (r + 1) * bh;                 // :413-415
        if (r < rows - 1) ye += oy;
        xs = std::max(0.0, xs);                   // :418-421
        ys = std::max(0.0, ys);
        xe = std::min((double)page_w, xe);
        ye = std::min((double)page_h, ye);
        PgTileInfo ti;
        std::memset(&ti, 0, sizeof(ti));
        ti.x_start = xs; ti.y_start = ys; ti.x_end = xe; ti.y_end = ye;
        ti.grid_rows = rows; ti.grid_cols = cols; ti.row = r + 1; ti.col = c + 1;
        ti.x0 = (int32_t)xs; ti.y0 = (int32_t)ys; ti.x1 = (int32_t)xe; ti.y1 = (int32_t)ye;  // :424-427
        const int sw = ti.x1 - ti.x0, sh = ti.y1 - ti.y0;
        if (sw <= 0 || sh <= 0) {
          delete plan;
          pg_set_error("invalid argument: empty tile (grid %dx%d on %dx%d)", rows, cols, page_w, page_h);
          return PG_ERR_INVALID;
        }
        // LetterBox geometry (ultralytics published behaviour)
        double rr = std::min((double)imgsz / (double)sh, (double)imgsz / (double)sw);
        if (!scaleup) rr = std::min(rr, 1.0);
        ti.new_w = (int32_t)py_round_half_even((double)sw * rr);
        ti.new_h = (int32_t)py_round_half_even((double)sh * rr);
        if (ti.new_w < 1 || ti.new_h < 1 || ti.new_w > imgsz || ti.new_h > imgsz) {
          delete plan;
          pg_set_error("unsupported: tile %dx%d letterboxes to %dx%d", sw, sh, ti.new_w, ti.new_h);
          return PG_ERR_UNSUPPORTED;
        }
        int dwi = imgsz - ti.new_w, dhi = imgsz - ti.new_h;
        if (auto_pad) { dwi %= stride; dhi %= stride; }
        const double dw = dwi / 2.0, dh = dhi / 2.0;
        const int top = (int)py_round_half_even(dh - 0.1), bottom = (int)py_round_half_even(dh + 0.1);
        const int left = (int)py_round_half_even(dw - 0.1), right = (int)py_round_half_even(dw + 0.1);
        ti.pad_l = left; ti.pad_t = top;
        ti.out_w = ti.new_w + left + right;
        ti.out_h = ti.new_h + top + bottom;
        if (ti.out_w % 2 != 0) {
          delete plan;
          pg_set_error("unsupported: odd output width %d", ti.out_w);
          return PG_ERR_UNSUPPORTED;
        }
        ti.out_offset = out_off;
        out_off += (int64_t)3 * ti.out_h * ti.out_w;

        TileDev td;
        std::memset(&td, 0, sizeof(td));
        td.x0 = ti.x0; td.y0 = ti.y0; td.src_w = sw; td.src_h = sh;
        td.new_w = ti.new_w; td.new_h = ti.new_h; td.pad_l = ti.pad_l; td.pad_t = ti.pad_t;
        td.out_w = ti.out_w; td.out_h = ti.out_h;
        td.xtab_off = (int32_t)plan->xtab_full.size();
        td.ytab_off = (int32_t)plan->ytab.size();
        td.row_skew = (CHN * ti.x0) & 15;
        td.row_bytes = (td.row_skew + CHN * sw + 15) & ~15;
        td.chunk_x0 = 0; td.chunk_w = ti.out_w;
        td.has_xpad = (ti.pad_l != 0 || ti.new_w != ti.out_w) ? 1 : 0;
        td.out_off = ti.out_offset;
        std::vector<PgCoef> cx((size_t)ti.out_w);  // per output column; s0 < 0 marks a pad column
        for (int x = 0; x < ti.out_w; ++x) {
          const int rx = x - ti.pad_l;
          if (rx < 0 || rx >= ti.new_w) {
            cx[x].s0 = -1;
            plan->xtab_full.push_back(make_uint2(0u, TL_PADMARK));
          } else {
            cx[x] = pg_resize_coef(sw, ti.new_w, rx, true);
            // when c1 == 0 the neighbour is multiplied by zero, so s1 never needs to be stored
            plan->xtab_full.push_back(make_uint2((uint32_t)(CHN * cx[x].s0), (uint32_t)cx[x].c0 | ((uint32_t)cx[x].c1 << 16)));
          }
        }
        for (int y = 0; y < ti.new_h; ++y) {
          const PgCoef cf = pg_resize_coef(sh, ti.new_h, y, false);
          plan->ytab.push_back(make_int4(cf.s0, cf.s1, cf.c0, cf.c1));
        }
        plan->info.push_back(ti);
        plan->tiles_full.push_back(td);

        // column chunks: the fewest equal chunks (multiples of 64 px) whose staged rows fit TL_MAX_ROW_BYTES
        auto chunk_span = [&](int c0, int cw, int* lo, int* hi) {  // source pixels [lo, hi) sampled by the chunk
          int l = INT32_MAX, h = -1;
          for (int x = c0; x < std::min(c0 + cw, (int)ti.out_w); ++x) {
            if (cx[x].s0 < 0) continue;
            l = std::min(l, cx[x].s0);
            h = std::max(h, cx[x].s0 + (cx[x].c1 ? 1 : 0));
          }
          if (h < 0) { l = 0; h = 0; }  // nothing but pad columns: stage one pixel
          *lo = l; *hi = h + 1;
        };
        int n_chunks = 1, chunk_w = ti.out_w;
        for (;; ++n_chunks) {
          chunk_w = n_chunks == 1 ? ti.out_w : ((ti.out_w + n_chunks - 1) / n_chunks + 63) / 64 * 64;
          int worst = 0;
          for (int c0 = 0; c0 < ti.out_w; c0 += chunk_w) {
            int lo, hi;
            chunk_span(c0, chunk_w, &lo, &hi);
            worst = std::max(worst, ((CHN * (ti.x0 + lo)) & 15) + CHN * (hi - lo) + 15);
          }
          if (worst <= max_row_bytes || chunk_w <= 64) break;
        }
        for (int c0 = 0; c0 < ti.out_w; c0 += chunk_w) {
          TileDev ch = td;
          int lo, hi;
          chunk_span(c0, chunk_w, &lo, &hi);
          ch.chunk_x0 = c0;
          ch.chunk_w = std::min(chunk_w, (int)ti.out_w - c0);
          ch.x0 = ti.x0 + lo;
          ch.src_w = hi - lo;
          ch.row_skew = (CHN * ch.x0) & 15;
          ch.row_bytes = (ch.row_skew + CHN * (hi - lo) + 15) & ~15;
          ch.xtab_off = (int32_t)plan->xtab.size();
          ch.has_xpad = 0;
          for (int x = c0; x < c0 + ch.chunk_w; ++x) {
            if (cx[x].s0 < 0) {
              ch.has_xpad = 1;
              plan->xtab.push_back(make_uint2(0u, TL_PADMARK));
            } else {
              plan->xtab.push_back(make_uint2((uint32_t)(CHN * (cx[x].s0 - lo)), (uint32_t)cx[x].c0 | ((uint32_t)cx[x].c1 << 16)));
            }
          }
          plan->max_row_bytes = std::max(plan->max_row_bytes, ch.row_bytes);
          plan->max_out_w = std::max(plan->max_out_w, ch.chunk_w);
          // work items of the chunk: bands of TL_BAND output rows, keyed by the page row their source starts at
          for (int oy0 = 0; oy0 < ti.out_h; oy0 += TL_BAND) {
            const int ry = std::min(std::max(oy0 - ti.pad_t, 0), ti.new_h - 1);
            const int key = ti.y0 + plan->ytab[(size_t)td.ytab_off + ry].x;
            plan->items.push_back(make_int4((int)plan->tiles.size(), oy0, std::min(TL_BAND, ti.out_h - oy0), key));
          }
          plan->tiles.push_back(ch);
        }
      }
    }
  }
  // One launch covers every grid: order the bands by the page row they read, so that the tiles of all
  // grids (and the 20 % overlaps inside a grid) that sample the same part of the page are in flight
  // together and share it through L2.  Stable: columns of one tile row stay adjacent.
  std::stable_sort(plan->items.begin(), plan->items.end(), [](const int4& l, const int4& r) { return l.w < r.w; });
  for (int4& it : plan->items) it.w = 0;
  plan->out_elems = out_off;
  *plan_out = plan;
  return PG_OK;
}

static void plan_free_device(PgTilePlan* p) {
  if (p->d_tiles) cudaFree(p->d_tiles);
  if (p->d_tiles_full) cudaFree(p->d_tiles_full);
  if (p->d_xtab) cudaFree(p->d_xtab);
  if (p->d_xtab_full) cudaFree(p->d_xtab_full);
  if (p->d_ytab) cudaFree(p->d_ytab);
  if (p->d_items) cudaFree(p->d_items);
  if (p->d_counters) cudaFree(p->d_counters);
  p->d_tiles = nullptr; p->d_xtab = nullptr; p->d_ytab = nullptr; p->d_items = nullptr; p->d_counters = nullptr;
  p->d_tiles_full = nullptr; p->d_xtab_full = nullptr;
  p->device = -1;
}

extern "C" void pg_tile_plan_destroy(PgTilePlan* plan) {
  if (!plan) return;
  plan_free_device(plan);
  delete plan;
}
extern "C" int32_t pg_tile_plan_num_tiles(const PgTilePlan* plan) { return plan ? (int32_t)plan->info.size() : 0; }
extern "C" int pg_tile_plan_tile(const PgTilePlan* plan, int32_t tile, PgTileInfo* info) {
  PG_REQUIRE(plan && info && tile >= 0 && tile < (int32_t)plan->info.size(), "tile index");
  *info = plan->info[tile];
  return PG_OK;
}
extern "C" int64_t pg_tile_plan_out_elems(const PgTilePlan* plan) { return plan ? plan->out_elems : 0; }
extern "C" int64_t pg_tile_plan_algorithmic_bytes(const PgTilePlan* plan) {
  return plan ? (int64_t)plan->channels * plan->page_w * plan->page_h + 2 * plan->out_elems : 0;
}

static int plan_upload(PgTilePlan* p, cudaStream_t s) {
  int dev = 0;
  PG_CUDA_TRY(cudaGetDevice(&dev));
  if (p->device == dev) return PG_OK;
  plan_free_device(p);
  PG_CUDA_TRY(cudaMalloc(&p->d_tiles, p->tiles.size() * sizeof(TileDev)));
  PG_CUDA_TRY(cudaMalloc(&p->d_xtab, p->xtab.size() * sizeof(uint2)));
  PG_CUDA_TRY(cudaMalloc(&p->d_tiles_full, p->tiles_full.size() * sizeof(TileDev)));
  PG_CUDA_TRY(cudaMalloc(&p->d_xtab_full, p->xtab_full.size() * sizeof(uint2)));
  PG_CUDA_TRY(cudaMalloc(&p->d_ytab, p->ytab.size() * sizeof(int4)));
  PG_CUDA_TRY(cudaMalloc(&p->d_items, p->items.size() * sizeof(int4)));
  PG_CUDA_TRY(cudaMalloc(&p->d_counters, 2 * TL_COUNTER_SLOTS * sizeof(unsigned long long)));
  PG_CUDA_TRY(cudaMemsetAsync(p->d_counters, 0, 2 * TL_COUNTER_SLOTS * sizeof(unsigned long long), s));
  PG_CUDA_TRY(cudaMemcpyAsync(p->d_tiles, p->tiles.data(), p->tiles.size() * sizeof(TileDev), cudaMemcpyHostToDevice, s));
  PG_CUDA_TRY(cudaMemcpyAsync(p->d_xtab, p->xtab.data(), p->xtab.size() * sizeof(uint2), cudaMemcpyHostToDevice, s));
  PG_CUDA_TRY(cudaMemcpyAsync(p->d_tiles_full, p->tiles_full.data(), p->tiles_full.size() * sizeof(TileDev), cudaMemcpyHostToDevice, s));
  PG_CUDA_TRY(cudaMemcpyAsync(p->d_xtab_full, p->xtab_full.data(), p->xtab_full.size() * sizeof(uint2), cudaMemcpyHostToDevice, s));
  PG_CUDA_TRY(cudaMemcpyAsync(p->d_ytab, p->ytab.data(), p->ytab.size() * sizeof(int4), cudaMemcpyHostToDevice, s));
  PG_CUDA_TRY(cudaMemcpyAsync(p->d_items, p->items.data(), p->items.size() * sizeof(int4), cudaMemcpyHostToDevice, s));
  PG_CUDA_TRY(cudaStreamSynchronize(s));  // host vectors may not be pinned; one-time cost
  p->device = dev;
  return PG_OK;
}

// ------------------------------------------------------------------------------------------
// device helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// the hot loops pass 32-bit shared-window addresses computed once (a generic->shared conversion per call
// costs an S2R and three ALU instructions per output row)
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) { mbar_arrive(smem_u32(bar)); }
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must surface as a trap (launch error), never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 1023u) == 0) {
      unsigned long long t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > 4000000000ull) __trap();  // 4 s
    }
  }
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) { mbar_wait(smem_u32(bar), parity); }
__device__ __forceinline__ int4 lds128(uint32_t addr) {
  int4 v;
  asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}

// 8 source bytes starting at byte offset `off` of a staged row -> horizontal pass of one pixel.
// lo = B0 G0 R0 B1, hi = G1 R1 . .   (BGR interleaved source, neighbour pixel 3 bytes on)
// `off` arrives packed (pack_xoff): word-aligned byte offset in the high half, bit shift in the low
// five bits — one LEA.HI forms the address and the wrapping funnel shift takes the register as is.
// A staged row is < 64 KB (two rows x two stages must fit 227 KB), so the offset fits 16 bits.
__device__ __forceinline__ uint32_t pack_xoff(uint32_t off) {
  // bit 5 (ignored by the wrapping shift): the pixel pair's last byte lies in the NEXT word — only for off % 4 == 3
  // (BGR: bytes off .. off+5; grey: off, off+1), so that word's load is predicated and three lanes in four skip it
  return ((off & ~3u) << 16) | ((off & 3u) * 8u) | ((off & 3u) == 3u ? 32u : 0u);
}
#ifndef PG_TILER_TAIL_PRED
#define PG_TILER_TAIL_PRED 1  // 0: every lane loads the tail word (A/B builds)
#endif
// The aligned 8-byte windows (lo, hi) = bytes off .. off+3 and off+4 .. off+7 of both staged rows of one pixel.  The
// word behind the second is loaded only by the lanes whose pack_xoff flag is set, INTO the register of the first
// word once `lo` has been formed — defined on every path, so no register has to be initialised for the other lanes
// (their `hi` takes no byte from it: shift < 24).
__device__ __forceinline__ void lds_windows_rows(uint32_t a0, uint32_t a1, uint32_t off, uint32_t& lo0, uint32_t& hi0,
                                                 uint32_t& lo1, uint32_t& hi1) {
#if PG_TILER_TAIL_PRED
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b32 f, x0, x1, y0, y1;\n\t"
      "and.b32 f, %6, 32;\n\t"
      "setp.ne.u32 p, f, 0;\n\t"
      "ld.shared.u32 x0, [%4];\n\t"
      "ld.shared.u32 y0, [%5];\n\t"
      "ld.shared.u32 x1, [%4+4];\n\t"
      "ld.shared.u32 y1, [%5+4];\n\t"
      "shf.r.wrap.b32 %0, x0, x1, %6;\n\t"
      "shf.r.wrap.b32 %2, y0, y1, %6;\n\t"
      "@p ld.shared.u32 x0, [%4+8];\n\t"
      "@p ld.shared.u32 y0, [%5+8];\n\t"
      "shf.r.wrap.b32 %1, x1, x0, %6;\n\t"
      "shf.r.wrap.b32 %3, y1, y0, %6;\n\t}"
      : "=&r"(lo0), "=&r"(hi0), "=&r"(lo1), "=&r"(hi1)
      : "r"(a0), "r"(a1), "r"(off));
#else
  const uint32_t p0 = lds32(a0), p1 = lds32(a0 + 4), p2 = lds32(a0 + 8), q0 = lds32(a1), q1 = lds32(a1 + 4), q2 = lds32(a1 + 8);
  lo0 = __funnelshift_r(p0, p1, off); hi0 = __funnelshift_r(p1, p2, off);
  lo1 = __funnelshift_r(q0, q1, off); hi1 = __funnelshift_r(q1, q2, off);
#endif
}
// one-channel pages: the two neighbouring source bytes of a pixel at byte offset `off` of a staged row
__device__ __forceinline__ uint32_t hpass_smem_grey(uint32_t row_addr, uint32_t off, uint32_t coef) {
  const uint32_t a = row_addr + (off >> 16);
  const uint32_t w0 = lds32(a), w1 = lds32(a + 4);
  return __dp2a_lo(coef, __funnelshift_r(w0, w1, off), 0u);  // a0*S0 + a1*S1
}

// both source rows of one pixel at once (the predicate of the tail word is shared)
__device__ __forceinline__ void hpass_smem_rows(uint32_t row0, uint32_t row1, uint32_t off, uint32_t coef, uint32_t (&t)[3],
                                                uint32_t (&u)[3]) {
  uint32_t lo0, hi0, lo1, hi1;
  lds_windows_rows(row0 + (off >> 16), row1 + (off >> 16), off, lo0, hi0, lo1, hi1);
  {
    const uint32_t bg = __byte_perm(lo0, hi0, 0x4130), rr = __byte_perm(lo0, hi0, 0x0052);  // B0 B1 G0 G1 | R0 R1 . .
    t[0] = __dp2a_lo(coef, bg, 0u); t[1] = __dp2a_hi(coef, bg, 0u); t[2] = __dp2a_lo(coef, rr, 0u);
  }
  {
    const uint32_t bg = __byte_perm(lo1, hi1, 0x4130), rr = __byte_perm(lo1, hi1, 0x0052);
    u[0] = __dp2a_lo(coef, bg, 0u); u[1] = __dp2a_hi(coef, bg, 0u); u[2] = __dp2a_lo(coef, rr, 0u);
  }
}
__device__ __forceinline__ uint32_t pack_unit_half2(uint32_t va, uint32_t vb) {
  // fp16(v/255): v*(1/255) in fp32 then one RN conversion matches fp16(fp32(v)/255) for all 256 v
  const float k = 1.0f / 255.0f;
  __half2 h = __floats2half2_rn((float)va * k, (float)vb * k);
  return *reinterpret_cast<uint32_t*>(&h);
}

// Second half of pg_vpass for one pixel pair of one channel, 16 bits to a lane.  a[j] = b0*(h0>>4) +
// (2 << 16) and b[j] = b1*(h1>>4) are the row products of pixel j; the sum of their high halves is
// 4v+r (< 1024), so v = sum >> 2 lands in byte 1 of each lane after a multiply by 64.  The two bytes
// are then read as fp16 SUBNORMALS (0x00vv = v * 2^-24, exact) and scaled by 2^24/255 split in two
// fp16 constants: fma(x, 64608, x*1185) equals fp16(fp32(v) * (1/255)) for all 256 v (checked
// exhaustively on the host: tests/test_cabi_host.py::test_half2_unit_scale_exhaustive).  The FMA pipe
// is idle in this kernel while the ALU pipe is its second limiter, hence the trade.
__device__ __forceinline__ uint32_t vpass_pair_half2(const uint32_t (&a)[2], const uint32_t (&b)[2]) {
  const uint32_t s = __byte_perm(a[0], a[1], 0x7632) + __byte_perm(b[0], b[1], 0x7632);
  const uint32_t v = __byte_perm(s * 64u, 0u, 0x4341);
  const __half2 x = *reinterpret_cast<const __half2*>(&v);
  const __half2 c_hi = __halves2half2(__ushort_as_half(0x7be3), __ushort_as_half(0x7be3));  // 64608
  const __half2 c_lo = __halves2half2(__ushort_as_half(0x64a1), __ushort_as_half(0x64a1));  // 1185
  const __half2 r = __hfma2(x, c_hi, __hmul2(x, c_lo));
  return *reinterpret_cast<const uint32_t*>(&r);
}

// ------------------------------------------------------------------------------------------
// K1 persistent pipeline kernel
// consumer warps per CTA: template parameter CW (8, or 4 — half the threads and twice the pixels per thread and row, so
// the per-row hand-shake weighs half as much; knob PG_TILER_CW).  A CTA has CW + 1 warps (the last is the producer).
// Measured on B200 (cfg3, 64 pages): ring depth 3 -> 4 CTAs/SM, 2.22 ms (0.97 of HBM peak) vs depth 4 ->
// 3 CTAs/SM, 2.26 ms; claiming <= 4 bands per CTA beats 8/16/64 (2.26 / 2.28 / 2.39 ms at depth 4).
constexpr int TL_STAGES_DEFAULT = 3;         // shared-memory ring depth (template parameter TL_STAGES)
constexpr int TL_ITEMS_PER_CTA = 4;           // bands of TL_BAND rows a CTA claims before retiring
#define TL_PAIR_STRIDE (CW * 64)              // pixels covered by all consumer warps per iteration (inside templates on CW)

// one page of a heterogeneous batch (pages of different sizes in one launch)
struct PageDesc {
  const uint8_t* src;  // page pixels
  __half* out;         // this page's tile buffer
  int64_t pitch;
  long long item_off;  // first global work item of this page
  int32_t item_base;   // where this page's plan starts in the concatenated item table
  int32_t tile_base;   // ... and in the concatenated tile table
};

struct TilerArgs {
  const PageDesc* page_desc;  // non-null: heterogeneous batch (pages/out/pitch/strides below unused)
  int32_t n_pages;
  const uint8_t* pages;
  __half* out;
  const TileDev* tiles;
  const uint2* xtab;
  const int4* ytab;
  const int4* items;
  int64_t pitch, page_stride, out_page_stride;
  int32_t items_per_page;
  int64_t total_items;
  int32_t row_stride;     // shared-memory bytes per staged row
  int32_t items_per_cta;  // work items a CTA may claim before it retires
  int32_t channels;       // bytes per source pixel (3: BGR, 1: grey plane)
  unsigned long long* counters;  // [0] next item, [1] CTAs finished (self-resetting)
};

// Stage message, written by the producer before it arms the stage's full barrier:
//   x = tile (-1: no more work), y = page, z = output row | TL_MSG_PADROW, w = b0 | b1 << 16
constexpr int TL_MSG_PADROW = 1 << 30;

// One output row.  Per iteration a warp covers 64 consecutive pixels and lane L computes pixels L (j = 0) and
// 32 + L (j = 1): ADJACENT lanes read ADJACENT source pixels, so one LDS instruction of the warp spans half as
// many shared-memory words as with a pixel pair per lane (the kernel's l1tex pipe was 92 % busy with 2.3
// wavefronts per LDS).  The half2 pairs for the stores are formed by one exchange with the neighbouring lane:
// even lanes store pixels (L, L+1), odd lanes store (32+L-1, 32+L).  XPAD: the chunk has 114-valued columns.
template <int ITER, bool XPAD, int CHN, int CW>
__device__ __forceinline__ void tiler_row(uint32_t row0, uint32_t row1, const uint32_t (&xoff)[ITER][2],
                                          const uint32_t (&coef)[ITER][2], uint32_t b0, uint32_t b1,
                                          uint32_t* pr, uint32_t* pg, uint32_t* pb, int out_w, int warp_px, int store_px,
                                          uint32_t pair_sel) {
#pragma unroll
  for (int i = 0; i < ITER; ++i) {
    if (i * TL_PAIR_STRIDE + warp_px < out_w) {  // warp-uniform: the exchange below needs every lane
      if (CHN == 1) {
        // one grey plane in, the same value to the three output planes (what cv2.imread's replicated channels
        // would give through the three-channel path, at a third of the arithmetic and of the page bytes)
        uint32_t ay[2], by[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          if (XPAD && coef[i][j] == TL_PADMARK) {
            ay[j] = (4u * TL_PAD_VALUE + 2u) << 16;
            by[j] = 0u;
          } else {
            PG_DEV_ASSERT((xoff[i][j] >> 16) + 8u <= row1 - row0);
            ay[j] = b0 * (hpass_smem_grey(row0, xoff[i][j], coef[i][j]) >> 4) + 0x20000u;
            by[j] = b1 * (hpass_smem_grey(row1, xoff[i][j], coef[i][j]) >> 4);
          }
        }
        const uint32_t vy = vpass_pair_half2(ay, by);
        const uint32_t ny = __shfl_xor_sync(0xffffffffu, vy, 1);
        if (i * TL_PAIR_STRIDE + store_px < out_w) {
          const uint32_t pair = __byte_perm(vy, ny, pair_sel);
          __stcs(pr + i * (TL_PAIR_STRIDE / 2), pair);
          __stcs(pg + i * (TL_PAIR_STRIDE / 2), pair);
          __stcs(pb + i * (TL_PAIR_STRIDE / 2), pair);
        }
        continue;
      }
      // pg_vpass split in two: the 32-bit products of each source row (rounding +2 folded into the
      // row-0 product's addend), then both pixels of the pair finished 16 bits to a lane.
      uint32_t ab[2], ag[2], ar[2], bb[2], bg[2], br[2];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        if (XPAD && coef[i][j] == TL_PADMARK) {
          ab[j] = ag[j] = ar[j] = (4u * TL_PAD_VALUE + 2u) << 16;
          bb[j] = bg[j] = br[j] = 0u;
        } else {
          uint32_t t[3], u[3];
          PG_DEV_ASSERT((xoff[i][j] >> 16) + 12u <= row1 - row0);  // the three words stay inside the staged row
          hpass_smem_rows(row0, row1, xoff[i][j], coef[i][j], t, u);
          ab[j] = b0 * (t[0] >> 4) + 0x20000u;  bb[j] = b1 * (u[0] >> 4);
          ag[j] = b0 * (t[1] >> 4) + 0x20000u;  bg[j] = b1 * (u[1] >> 4);
          ar[j] = b0 * (t[2] >> 4) + 0x20000u;  br[j] = b1 * (u[2] >> 4);
        }
      }
      // BGR -> RGB planes
      // pr/pg/pb already point at this thread's first pixel pair of the row in each plane
      const uint32_t vr = vpass_pair_half2(ar, br), vg = vpass_pair_half2(ag, bg), vb = vpass_pair_half2(ab, bb);
      const uint32_t nr = __shfl_xor_sync(0xffffffffu, vr, 1), ng = __shfl_xor_sync(0xffffffffu, vg, 1),
                     nb = __shfl_xor_sync(0xffffffffu, vb, 1);
      if (i * TL_PAIR_STRIDE + store_px < out_w) {
        __stcs(pr + i * (TL_PAIR_STRIDE / 2), __byte_perm(vr, nr, pair_sel));
        __stcs(pg + i * (TL_PAIR_STRIDE / 2), __byte_perm(vg, ng, pair_sel));
        __stcs(pb + i * (TL_PAIR_STRIDE / 2), __byte_perm(vb, nb, pair_sel));
      }
    }
  }
}

template <int ITER, int TL_STAGES, int CHN, int CW>
__global__ void __launch_bounds__(32 * (CW + 1), (ITER <= 2 || CW <= 4) ? 4 : 2) tile_letterbox_kernel(const TilerArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t full_bar[TL_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[TL_STAGES];
  __shared__ __align__(16) int4 msg[TL_STAGES];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < TL_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], CW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const uint32_t stage_bytes = 2u * (uint32_t)a.row_stride;
  uint32_t stage = 0, phase = 0;

  if (warp == CW) {
    // ============ producer: one elected lane claims work items and issues the bulk copies ============
    if (lane == 0) {
      for (int n = 0; n < a.items_per_cta; ++n) {
        const long long it = (long long)atomicAdd(&a.counters[0], 1ull);
        if (it >= a.total_items) break;
        long long page;
        int4 item;
        int64_t pitch;
        const uint8_t* page_src;
        if (a.page_desc) {  // heterogeneous batch: find the page that owns work item `it`
          int lo = 0, hi = a.n_pages - 1;
          while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (a.page_desc[mid].item_off <= it) lo = mid; else hi = mid - 1;
          }
          page = lo;
          const PageDesc& pd = a.page_desc[lo];
          item = a.items[pd.item_base + (int)(it - pd.item_off)];
          item.x += pd.tile_base;
          pitch = pd.pitch;
          page_src = pd.src;
        } else {
          page = it / a.items_per_page;
          item = a.items[it - page * a.items_per_page];
          pitch = a.pitch;
          page_src = a.pages + page * a.page_stride;
        }
        const TileDev& t = a.tiles[item.x];
        const int pad_t = t.pad_t, new_h = t.new_h, ytab_off = t.ytab_off;
        const uint8_t* src = page_src + (int64_t)t.y0 * pitch + ((CHN * t.x0) & ~15);
        const uint32_t bytes = (uint32_t)t.row_bytes;
        PG_DEV_ASSERT(bytes <= (uint32_t)a.row_stride && (bytes & 15u) == 0u && item.x >= 0 && page >= 0 && page < a.n_pages &&
                      (int64_t)((CHN * t.x0) & ~15) + bytes <= pitch);
        for (int oy = item.y; oy < item.y + item.z; ++oy) {
          const int ry = oy - pad_t;
          const bool pad_row = ry < 0 || ry >= new_h;
          int4 yt = make_int4(0, 0, 0, 0);
          if (!pad_row) yt = a.ytab[ytab_off + ry];
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          msg[stage] = make_int4(item.x, (int)page, oy | (pad_row ? TL_MSG_PADROW : 0), yt.z | (yt.w << 16));
          if (pad_row) {
            mbar_arrive(&full_bar[stage]);  // nothing to stage: the message alone completes the phase
          } else {
            mbar_expect_tx(&full_bar[stage], 2u * bytes);
            uint8_t* dst = smem + stage * stage_bytes;
            bulk_g2s(dst, src + (int64_t)yt.x * pitch, bytes, &full_bar[stage]);
            bulk_g2s(dst + a.row_stride, src + (int64_t)yt.y * pitch, bytes, &full_bar[stage]);
          }
          if (++stage == TL_STAGES) { stage = 0; phase ^= 1u; }
        }
      }
      mbar_wait(&empty_bar[stage], phase ^ 1u);
      msg[stage] = make_int4(-1, 0, 0, 0);
      mbar_arrive(&full_bar[stage]);
      // the last CTA to retire re-arms the counters for the next launch on this plan
      __threadfence();
      const unsigned long long prev = atomicAdd(&a.counters[1], 1ull);
      if (prev == (unsigned long long)gridDim.x - 1ull) {
        a.counters[0] = 0ull;
        a.counters[1] = 0ull;
        __threadfence();
      }
    }
    return;
  }

  // ===================== consumers =====================
  uint32_t smem_base = smem_u32(smem);
  uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]), msg0 = smem_u32(&msg[0]);
  // opaque to the optimiser: otherwise it re-derives each address from SR_CgaCtaId on every output row
  asm volatile("" : "+r"(smem_base), "+r"(full0), "+r"(empty0), "+r"(msg0));
  const uint32_t pad_pair = pack_unit_half2(TL_PAD_VALUE, TL_PAD_VALUE);
  const int warp_px = warp * 64;                                              // first pixel of the warp's 64
  const int store_px = warp_px + ((lane & 1) ? 31 + lane : lane);             // first pixel of the pair this lane stores
  const uint32_t pair_sel = (lane & 1) ? 0x3276u : 0x5410u;                   // (nb.hi, own.hi) : (own.lo, nb.lo)
  uint32_t xoff[ITER][2], coef[ITER][2];
  int cached_tile = -1, cached_page = -1, out_w = 0, chunk_w = 0, chunk_x0 = 0, xpad = 0;
  int64_t plane_b = 0, out_off = 0;   // plane size in bytes
  uint32_t row_b = 0;                 // output row pitch in bytes
  uint8_t* tile_ptr = nullptr;        // this thread's first pixel pair of row 0, R plane

  while (true) {
    mbar_wait(full0 + stage * 8u, phase);
    const int4 m = lds128(msg0 + stage * 16u);
    if (m.x < 0) break;
    if (m.x != cached_tile) {
      cached_tile = m.x;
      const TileDev& t = a.tiles[m.x];
      out_w = t.out_w;
      chunk_w = t.chunk_w;
      chunk_x0 = t.chunk_x0;
      plane_b = (int64_t)t.out_h * out_w * 2;
      row_b = (uint32_t)out_w * 2u;
      out_off = t.out_off;
      cached_page = -1;
      xpad = t.has_xpad;
      const uint32_t skew = (uint32_t)t.row_skew;
      const int xtab_off = t.xtab_off;
#pragma unroll
      for (int i = 0; i < ITER; ++i) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int ox = i * TL_PAIR_STRIDE + warp_px + j * 32 + lane;
          uint2 e = make_uint2(0u, TL_PADMARK);
          if (ox < chunk_w) e = __ldg(&a.xtab[xtab_off + ox]);
          xoff[i][j] = pack_xoff(e.x + skew);
          coef[i][j] = e.y;
        }
      }
    }
    if (m.y != cached_page) {
      cached_page = m.y;
      __half* page_out = a.page_desc ? a.page_desc[m.y].out : a.out + (int64_t)m.y * a.out_page_stride;
      tile_ptr = reinterpret_cast<uint8_t*>(page_out + out_off) + (chunk_x0 + store_px) * 2;
    }
    const int oy = m.z & ~TL_MSG_PADROW;
    uint32_t* pr = reinterpret_cast<uint32_t*>(tile_ptr + (uint64_t)((uint32_t)oy * row_b));  // a plane is < 4 GB
    uint32_t* pg = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(pr) + plane_b);
    uint32_t* pb = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(pg) + plane_b);
    if (m.z & TL_MSG_PADROW) {
#pragma unroll
      for (int i = 0; i < ITER; ++i) {
        if (i * TL_PAIR_STRIDE + store_px < chunk_w) {
          __stcs(pr + i * (TL_PAIR_STRIDE / 2), pad_pair);
          __stcs(pg + i * (TL_PAIR_STRIDE / 2), pad_pair);
          __stcs(pb + i * (TL_PAIR_STRIDE / 2), pad_pair);
        }
      }
    } else {
      const uint32_t b0 = (uint32_t)m.w & 0xFFFFu, b1 = (uint32_t)m.w >> 16;
      const uint32_t row0 = smem_base + stage * stage_bytes;
      const uint32_t row1 = row0 + (uint32_t)a.row_stride;
      if (xpad) tiler_row<ITER, true, CHN, CW>(row0, row1, xoff, coef, b0, b1, pr, pg, pb, chunk_w, warp_px, store_px, pair_sel);
      else tiler_row<ITER, false, CHN, CW>(row0, row1, xoff, coef, b0, b1, pr, pg, pb, chunk_w, warp_px, store_px, pair_sel);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(empty0 + stage * 8u);
    if (++stage == TL_STAGES) { stage = 0; phase ^= 1u; }
  }
}

// ------------------------------------------------------------------------------------------
// validation twin: plain global loads, one thread per output pixel pair
__global__ void __launch_bounds__(256) tile_letterbox_direct_kernel(const TilerArgs a, int32_t n_tiles, int32_t n_pages) {
  const int tile = blockIdx.y % n_tiles, page = blockIdx.y / n_tiles;
  if (page >= n_pages) return;
  const TileDev t = a.tiles[tile];
  const int pairs_w = t.out_w >> 1;
  const int64_t npairs = (int64_t)pairs_w * t.out_h;
  __half* out_tile = a.out + (int64_t)page * a.out_page_stride + t.out_off;
  const int64_t plane = (int64_t)t.out_h * t.out_w;
  const int chn = a.channels;
  const uint8_t* src = a.pages + (int64_t)page * a.page_stride + (int64_t)t.y0 * a.pitch + chn * (int64_t)t.x0;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < npairs; p += (int64_t)gridDim.x * blockDim.x) {
    const int oy = (int)(p / pairs_w), ox = (int)(p - (int64_t)oy * pairs_w) * 2;
    const int ry = oy - t.pad_t;
    uint32_t v[3][2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const uint2 e = a.xtab[t.xtab_off + ox + j];
      if (ry < 0 || ry >= t.new_h || e.y == TL_PADMARK) {
        v[0][j] = v[1][j] = v[2][j] = TL_PAD_VALUE;
      } else {
        const int4 yt = a.ytab[t.ytab_off + ry];
        const uint32_t a0 = e.y & 0xFFFFu, a1 = e.y >> 16;
        const uint8_t* r0 = src + (int64_t)yt.x * a.pitch + e.x;
        const uint8_t* r1 = src + (int64_t)yt.y * a.pitch + e.x;
        const int dx = a1 ? chn : 0;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const int cc = chn == 3 ? c : 0;  // a grey plane feeds all three output planes
          const uint32_t h0 = pg_hpass(r0[cc], r0[cc + dx], a0, a1);
          const uint32_t h1 = pg_hpass(r1[cc], r1[cc + dx], a0, a1);
          v[c][j] = pg_vpass(h0, h1, (uint32_t)yt.z, (uint32_t)yt.w);
        }
      }
    }
    uint32_t* o = reinterpret_cast<uint32_t*>(out_tile + (int64_t)oy * t.out_w);
    o[ox >> 1] = pack_unit_half2(v[2][0], v[2][1]);
    o[(plane + ox) >> 1] = pack_unit_half2(v[1][0], v[1][1]);
    o[(2 * plane + ox) >> 1] = pack_unit_half2(v[0][0], v[0][1]);
  }
}

static int check_pages_layout(const PgTilePlan* plan, const uint8_t* pages, int32_t n_pages, int64_t pitch,
                              int64_t page_stride, const void* out, int64_t out_page_stride) {
  PG_REQUIRE(plan != nullptr, "plan");
  PG_REQUIRE(pages != nullptr && out != nullptr, "null device pointer");
  PG_REQUIRE(n_pages >= 0, "n_pages");
  PG_REQUIRE(pitch >= (int64_t)plan->channels * plan->page_w && pitch % 16 == 0,
             "pitch must be >= channels*W and a multiple of 16");
  PG_REQUIRE(page_stride >= pitch * plan->page_h && page_stride % 16 == 0, "page_stride");
  PG_REQUIRE(((uintptr_t)pages & 15) == 0, "pages must be 16-byte aligned");
  PG_REQUIRE(out_page_stride >= plan->out_elems && out_page_stride % 2 == 0, "out_page_stride");
  PG_REQUIRE(((uintptr_t)out & 3) == 0, "out must be 4-byte aligned");
  return PG_OK;
}

static TilerArgs make_args(const PgTilePlan* plan, const uint8_t* pages, int32_t n_pages, int64_t pitch,
                           int64_t page_stride, void* out, int64_t out_page_stride) {
  TilerArgs a;
  a.page_desc = nullptr;
  a.n_pages = n_pages;
  a.pages = pages;
  a.out = reinterpret_cast<__half*>(out);
  a.tiles = plan->d_tiles;
  a.xtab = plan->d_xtab;
  a.ytab = plan->d_ytab;
  a.items = plan->d_items;
  a.pitch = pitch;
  a.page_stride = page_stride;
  a.out_page_stride = out_page_stride;
  a.items_per_page = (int32_t)plan->items.size();
  a.total_items = (int64_t)plan->items.size() * n_pages;
  a.row_stride = (plan->max_row_bytes + 16 + 127) & ~127;
  a.items_per_cta = 1;
  a.channels = plan->channels;
  a.counters = plan->d_counters;  // the pipeline launch picks its own slot (counter_slot)
  return a;
}

template <int ITER, int TL_STAGES, int CHN, int CW>
static int launch_pipeline_c(TilerArgs a, cudaStream_t s) {
  constexpr int TL_THREADS = 32 * (CW + 1);
  const size_t smem = (size_t)TL_STAGES * 2 * a.row_stride;
  int dev = 0, sms = 0, max_smem = 0;
  PG_CUDA_TRY(cudaGetDevice(&dev));
  PG_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  PG_CUDA_TRY(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  if (smem + 1024 > (size_t)max_smem) {
    pg_set_error("unsupported: tile rows of %d bytes need %zu B of shared memory (max %d)", a.row_stride, smem, max_smem);
    return PG_ERR_UNSUPPORTED;
  }
  auto kernel = tile_letterbox_kernel<ITER, TL_STAGES, CHN, CW>;
  PG_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  PG_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, TL_THREADS, smem));
  if (per_sm < 1) {
    pg_set_error("unsupported: tiler kernel does not fit on an SM");
    return PG_ERR_UNSUPPORTED;
  }
  // Bounded-lifetime CTAs: each claims up to items_per_cta bands of 16 rows from the shared counter
  // and retires, so kernels queued on a higher-priority stream (the box stages) can take over SM
  // slots while the tiler is still streaming.  Small launches get 1 item per CTA to fill the chip.
  const int64_t slots = (int64_t)sms * per_sm;
  int64_t ipc = a.total_items / (slots * 4);
  int64_t ipc_max = TL_ITEMS_PER_CTA;
  if (const char* e = getenv("PG_TILER_IPC")) ipc_max = atoi(e) > 0 ? atoi(e) : ipc_max;  // tuning knob
  ipc = ipc < 1 ? 1 : (ipc > ipc_max ? ipc_max : ipc);
  a.items_per_cta = (int32_t)ipc;
  const int64_t grid = (a.total_items + ipc - 1) / ipc;
  if (grid < 1) return PG_OK;
  if (grid > 0x7fffffffll) {
    pg_set_error("unsupported: %lld work items in one launch", (long long)a.total_items);
    return PG_ERR_UNSUPPORTED;
  }
  kernel<<<(unsigned)grid, TL_THREADS, smem, s>>>(a);
  PG_LAUNCH_CHECK();
  return PG_OK;
}

template <int ITER, int TL_STAGES, int CW>
static int launch_pipeline(const TilerArgs& a, cudaStream_t s) {
  return a.channels == 1 ? launch_pipeline_c<ITER, TL_STAGES, 1, CW>(a, s) : launch_pipeline_c<ITER, TL_STAGES, 3, CW>(a, s);
}

static int dispatch_pipeline(const TilerArgs& a, int max_out_w, cudaStream_t s);

extern "C" int pg_tile_letterbox(PgTilePlan* plan, const uint8_t* pages, int32_t n_pages, int64_t pitch,
                                 int64_t page_stride, void* out_f16, int64_t out_page_stride, void* stream) {
  int rc = check_pages_layout(plan, pages, n_pages, pitch, page_stride, out_f16, out_page_stride);
  if (rc != PG_OK) return rc;
  if (n_pages == 0) return PG_OK;
  cudaStream_t s = (cudaStream_t)stream;
  rc = plan_upload(plan, s);
  if (rc != PG_OK) return rc;
  TilerArgs a = make_args(plan, pages, n_pages, pitch, page_stride, out_f16, out_page_stride);
  a.counters = counter_slot(plan->d_counters, plan->launch_seq, s);
  return dispatch_pipeline(a, plan->max_out_w, s);
}

template <int CW>
static int dispatch_stages(const TilerArgs& a, int iters, int stages, cudaStream_t s) {
  if (stages == 2) {
    if (iters <= 1) return launch_pipeline<1, 2, CW>(a, s);
    if (iters <= 2) return launch_pipeline<2, 2, CW>(a, s);
    return launch_pipeline<4, 2, CW>(a, s);
  } else if (stages == 3) {
    if (iters <= 1) return launch_pipeline<1, 3, CW>(a, s);
    if (iters <= 2) return launch_pipeline<2, 3, CW>(a, s);
    return launch_pipeline<4, 3, CW>(a, s);
  }
  if (iters <= 1) return launch_pipeline<1, 4, CW>(a, s);
  if (iters <= 2) return launch_pipeline<2, 4, CW>(a, s);
  return launch_pipeline<4, 4, CW>(a, s);
}

static int dispatch_pipeline(const TilerArgs& a, int max_out_w, cudaStream_t s) {
  int stages = TL_STAGES_DEFAULT;
  if (const char* e = getenv("PG_TILER_STAGES")) stages = atoi(e);  // tuning knob: 3 (4 CTAs/SM) or 4 (3 CTAs/SM)
  int cw = 8;
  if (const char* e = getenv("PG_TILER_CW")) cw = atoi(e) == 4 ? 4 : 8;  // tuning knob: consumer warps per CTA
  // very wide tiles (a 1x1 grid on a > 12k px page): fall back to a 2-deep ring so the rows still fit
  int dev = 0, max_smem = 0;
  if (cudaGetDevice(&dev) == cudaSuccess &&
      cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) == cudaSuccess &&
      (size_t)stages * 2 * a.row_stride + 1024 > (size_t)max_smem)
    stages = 2;
  if (stages < 2 || stages > 4) stages = TL_STAGES_DEFAULT;
  if (cw == 4 && max_out_w <= 4 * 4 * 64) return dispatch_stages<4>(a, (max_out_w + 4 * 64 - 1) / (4 * 64), stages, s);
  if (max_out_w <= 4 * 8 * 64) return dispatch_stages<8>(a, (max_out_w + 8 * 64 - 1) / (8 * 64), stages, s);
  pg_set_error("unsupported: output width %d > %d", max_out_w, 4 * 8 * 64);
  return PG_ERR_UNSUPPORTED;
}

extern "C" int pg_tile_letterbox_direct(PgTilePlan* plan, const uint8_t* pages, int32_t n_pages, int64_t pitch,
                                        int64_t page_stride, void* out_f16, int64_t out_page_stride, void* stream) {
  int rc = check_pages_layout(plan, pages, n_pages, pitch, page_stride, out_f16, out_page_stride);
  if (rc != PG_OK) return rc;
  if (n_pages == 0) return PG_OK;
  cudaStream_t s = (cudaStream_t)stream;
  rc = plan_upload(plan, s);
  if (rc != PG_OK) return rc;
  TilerArgs a = make_args(plan, pages, n_pages, pitch, page_stride, out_f16, out_page_stride);
  a.tiles = plan->d_tiles_full;  // whole tiles and their own x tables: independent of the chunking
  a.xtab = plan->d_xtab_full;
  const int n_tiles = (int)plan->tiles_full.size();
  const int64_t gy = (int64_t)n_tiles * n_pages;
  PG_REQUIRE(gy <= 65535, "direct kernel: n_tiles*n_pages must be <= 65535");
  dim3 grid(64, (unsigned)gy);
  tile_letterbox_direct_kernel<<<grid, 256, 0, s>>>(a, n_tiles, n_pages);
  PG_LAUNCH_CHECK();
  return PG_OK;
}

// ------------------------------------------------------------------------------------------
// heterogeneous batches: pages of different sizes (one plan per size) tiled by ONE launch.
// The batch concatenates the plans' tables; the kernel maps a global work item to its page by a
// binary search over per-page item offsets (done by the producer lane, once per 16-row band).
struct PgTileBatch {
  std::vector<int32_t> page_plan;
  std::vector<int32_t> plan_w, plan_h, plan_tile_base, plan_item_base, plan_item_count;
  int32_t channels = 3;
  std::vector<int64_t> plan_out_elems;
  std::vector<TileDev> tiles;
  std::vector<uint2> xtab;
  std::vector<int4> ytab;
  std::vector<int4> items;
  std::vector<PageDesc> desc;
  int64_t total_items = 0, alg_bytes = 0;
  int32_t max_row_bytes = 0, max_out_w = 0;
  int device = -1;
  bool bound = false;
  TileDev* d_tiles = nullptr;
  uint2* d_xtab = nullptr;
  int4* d_ytab = nullptr;
  int4* d_items = nullptr;
  PageDesc* d_desc = nullptr;
  unsigned long long* d_counters = nullptr;
  std::atomic<uint32_t> launch_seq{0};
};

static void batch_free_device(PgTileBatch* b) {
  cudaFree(b->d_tiles); cudaFree(b->d_xtab); cudaFree(b->d_ytab); cudaFree(b->d_items); cudaFree(b->d_desc);
  cudaFree(b->d_counters);
  b->d_tiles = nullptr; b->d_xtab = nullptr; b->d_ytab = nullptr; b->d_items = nullptr; b->d_desc = nullptr;
  b->d_counters = nullptr;
  b->device = -1;
  b->bound = false;
}

extern "C" int pg_tile_batch_create(const PgTilePlan* const* plans, int32_t n_plans, const int32_t* page_plan,
                                    int32_t n_pages, PgTileBatch** out) {
  PG_REQUIRE(plans && page_plan && out && n_plans > 0 && n_pages > 0, "batch arguments");
  auto* b = new PgTileBatch();
  for (int k = 0; k < n_plans; ++k) {
    const PgTilePlan* p = plans[k];
    if (!p) {
      delete b;
      pg_set_error("invalid argument: plan %d is null", k);
      return PG_ERR_INVALID;
    }
    if (k == 0) b->channels = p->channels;
    if (p->channels != b->channels) {
      delete b;
      pg_set_error("invalid argument: the plans of a batch must share the channel count");
      return PG_ERR_INVALID;
    }
    const int32_t xb = (int32_t)b->xtab.size(), yb = (int32_t)b->ytab.size();
    b->plan_tile_base.push_back((int32_t)b->tiles.size());
    b->plan_item_base.push_back((int32_t)b->items.size());
    b->plan_item_count.push_back((int32_t)p->items.size());
    b->plan_w.push_back(p->page_w);
    b->plan_h.push_back(p->page_h);
    b->plan_out_elems.push_back(p->out_elems);
    for (TileDev t : p->tiles) {
      t.xtab_off += xb;
      t.ytab_off += yb;
      b->tiles.push_back(t);
    }
    b->xtab.insert(b->xtab.end(), p->xtab.begin(), p->xtab.end());
    b->ytab.insert(b->ytab.end(), p->ytab.begin(), p->ytab.end());
    b->items.insert(b->items.end(), p->items.begin(), p->items.end());
    b->max_row_bytes = std::max(b->max_row_bytes, p->max_row_bytes);
    b->max_out_w = std::max(b->max_out_w, p->max_out_w);
  }
  b->page_plan.assign(page_plan, page_plan + n_pages);
  long long off = 0;
  for (int i = 0; i < n_pages; ++i) {
    const int k = page_plan[i];
    if (k < 0 || k >= n_plans) {
      delete b;
      pg_set_error("invalid argument: page %d refers to plan %d", i, k);
      return PG_ERR_INVALID;
    }
    PageDesc d;
    d.src = nullptr; d.out = nullptr; d.pitch = 0;
    d.item_off = off;
    d.item_base = b->plan_item_base[k];
    d.tile_base = b->plan_tile_base[k];
    b->desc.push_back(d);
    off += b->plan_item_count[k];
    b->alg_bytes += (int64_t)b->channels * b->plan_w[k] * b->plan_h[k] + 2 * b->plan_out_elems[k];
  }
  b->total_items = off;
  *out = b;
  return PG_OK;
}

extern "C" void pg_tile_batch_destroy(PgTileBatch* b) {
  if (!b) return;
  batch_free_device(b);
  delete b;
}
extern "C" int64_t pg_tile_batch_algorithmic_bytes(const PgTileBatch* b) { return b ? b->alg_bytes : 0; }

extern "C" int pg_tile_batch_bind(PgTileBatch* b, const uint8_t* const* page_ptrs, const int64_t* pitches,
                                  void* const* out_ptrs, void* stream) {
  PG_REQUIRE(b && page_ptrs && pitches && out_ptrs, "batch bind arguments");
  cudaStream_t s = (cudaStream_t)stream;
  int dev = 0;
  PG_CUDA_TRY(cudaGetDevice(&dev));
  const size_t n = b->desc.size();
  for (size_t i = 0; i < n; ++i) {
    const int k = b->page_plan[i];
    PG_REQUIRE(page_ptrs[i] && out_ptrs[i], "null page / output pointer");
    PG_REQUIRE(((uintptr_t)page_ptrs[i] & 15) == 0 && pitches[i] % 16 == 0 && pitches[i] >= (int64_t)b->channels * b->plan_w[k],
               "each page must be 16-byte aligned with pitch >= channels*W and a multiple of 16");
    PG_REQUIRE(((uintptr_t)out_ptrs[i] & 3) == 0, "outputs must be 4-byte aligned");
    b->desc[i].src = page_ptrs[i];
    b->desc[i].out = reinterpret_cast<__half*>(out_ptrs[i]);
    b->desc[i].pitch = pitches[i];
  }
  if (b->device != dev) {
    batch_free_device(b);
    PG_CUDA_TRY(cudaMalloc(&b->d_tiles, b->tiles.size() * sizeof(TileDev)));
    PG_CUDA_TRY(cudaMalloc(&b->d_xtab, b->xtab.size() * sizeof(uint2)));
    PG_CUDA_TRY(cudaMalloc(&b->d_ytab, b->ytab.size() * sizeof(int4)));
    PG_CUDA_TRY(cudaMalloc(&b->d_items, b->items.size() * sizeof(int4)));
    PG_CUDA_TRY(cudaMalloc(&b->d_desc, n * sizeof(PageDesc)));
    PG_CUDA_TRY(cudaMalloc(&b->d_counters, 2 * TL_COUNTER_SLOTS * sizeof(unsigned long long)));
    PG_CUDA_TRY(cudaMemsetAsync(b->d_counters, 0, 2 * TL_COUNTER_SLOTS * sizeof(unsigned long long), s));
    PG_CUDA_TRY(cudaMemcpyAsync(b->d_tiles, b->tiles.data(), b->tiles.size() * sizeof(TileDev), cudaMemcpyHostToDevice, s));
    PG_CUDA_TRY(cudaMemcpyAsync(b->d_xtab, b->xtab.data(), b->xtab.size() * sizeof(uint2), cudaMemcpyHostToDevice, s));
    PG_CUDA_TRY(cudaMemcpyAsync(b->d_ytab, b->ytab.data(), b->ytab.size() * sizeof(int4), cudaMemcpyHostToDevice, s));
    PG_CUDA_TRY(cudaMemcpyAsync(b->d_items, b->items.data(), b->items.size() * sizeof(int4), cudaMemcpyHostToDevice, s));
    b->device = dev;
  }
  PG_CUDA_TRY(cudaMemcpyAsync(b->d_desc, b->desc.data(), n * sizeof(PageDesc), cudaMemcpyHostToDevice, s));
  PG_CUDA_TRY(cudaStreamSynchronize(s));  // host vectors are pageable; binding is a one-off per buffer set
  b->bound = true;
  return PG_OK;
}

extern "C" int pg_tile_letterbox_batch(PgTileBatch* b, void* stream) {
  PG_REQUIRE(b != nullptr, "batch");
  if (!b->bound) {
    pg_set_error("invalid argument: pg_tile_batch_bind must be called before pg_tile_letterbox_batch");
    return PG_ERR_INVALID;
  }
  TilerArgs a;
  a.page_desc = b->d_desc;
  a.n_pages = (int32_t)b->desc.size();
  a.pages = nullptr;
  a.out = nullptr;
  a.tiles = b->d_tiles;
  a.xtab = b->d_xtab;
  a.ytab = b->d_ytab;
  a.items = b->d_items;
  a.pitch = a.page_stride = a.out_page_stride = 0;
  a.items_per_page = 1;
  a.total_items = b->total_items;
  a.row_stride = (b->max_row_bytes + 16 + 127) & ~127;
  a.items_per_cta = 1;
  a.channels = b->channels;
  a.counters = counter_slot(b->d_counters, b->launch_seq, (cudaStream_t)stream);
  return dispatch_pipeline(a, b->max_out_w, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------
// synthetic pages (bench input generator; device-resident, counter-based)
__device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}

__global__ void synth_pages_kernel(uint8_t* pages, int32_t n_pages, int32_t w, int32_t h, int64_t pitch,
                                   int64_t page_stride, uint64_t seed0, int64_t first_page) {
  // one thread per 4 bytes of a row (pitch is a multiple of 16)
  const int64_t words_per_row = pitch >> 2;
  const int64_t total = (int64_t)n_pages * h * words_per_row;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / words_per_row;
    const int wx = (int)(i - row * words_per_row);
    const int page = (int)(row / h), y = (int)(row - (int64_t)page * h);
    const uint32_t pseed = mix32((uint32_t)(seed0 + (uint64_t)(first_page + page)) * 0x9E3779B9u + 0x85EBCA6Bu);
    uint32_t word = 0;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int byte = wx * 4 + b;
      const int x = byte / 3;
      uint32_t v = 0;
      if (x < w) {
        const uint32_t hp = mix32(pseed ^ mix32((uint32_t)y * 0x01000193u + (uint32_t)x));
        // text-line bands of 24 px with 9 px leading; ink probability higher inside a line
        const bool in_line = (y % 33) < 24;
        const bool ink = (hp & 0xFFu) < (in_line ? 70u : 4u);
        const uint32_t hc = mix32(hp + (uint32_t)(byte - 3 * x) + 1u);
        v = ink ? 20u + (hc & 31u) : 215u + (hc & 31u);
      }
      word |= v << (8 * b);
    }
    *reinterpret_cast<uint32_t*>(pages + (int64_t)page * page_stride + (int64_t)y * pitch + (int64_t)wx * 4) = word;
  }
}

extern "C" int pg_synth_pages(uint8_t* pages, int32_t n_pages, int32_t page_w, int32_t page_h, int64_t pitch,
                              int64_t page_stride, uint64_t seed0, int64_t first_page, void* stream) {
  PG_REQUIRE(pages != nullptr && n_pages >= 0 && page_w > 0 && page_h > 0, "pages");
  PG_REQUIRE(pitch >= (int64_t)3 * page_w && pitch % 16 == 0, "pitch");
  PG_REQUIRE(page_stride >= pitch * page_h && page_stride % 16 == 0, "page_stride");
  if (n_pages == 0) return PG_OK;
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  synth_pages_kernel<<<sms * 16, 256, 0, (cudaStream_t)stream>>>(pages, n_pages, page_w, page_h, pitch, page_stride,
                                                                  seed0, first_page);
  PG_LAUNCH_CHECK();
  return PG_OK;
}

// ------------------------------------------------------------------------------------------
// test hooks (host evaluations of pg_math.h)
extern "C" double pg_hostcheck_iou(const double* a, const double* b) {
  return pg_iou(a[0], a[1], a[2], a[3], pg_box_area(a[0], a[1], a[2], a[3]), b[0], b[1], b[2], b[3],
                pg_box_area(b[0], b[1], b[2], b[3]));
}
extern "C" int32_t pg_hostcheck_iou_gt(const double* a, const double* b, double thr) {
  return pg_iou_gt(a[0], a[1], a[2], a[3], pg_box_area(a[0], a[1], a[2], a[3]), b[0], b[1], b[2], b[3],
                   pg_box_area(b[0], b[1], b[2], b[3]), thr) ? 1 : 0;
}
extern "C" int32_t pg_hostcheck_iou_gt_f32(const float* a, const float* b, double thr) {
  return pg_iou_gt_f32(a[0], a[1], a[2], a[3], pg_box_area_f32(a[0], a[1], a[2], a[3]), b[0], b[1], b[2], b[3],
                       pg_box_area_f32(b[0], b[1], b[2], b[3]), thr) ? 1 : 0;
}
extern "C" int32_t pg_hostcheck_edge_touch(const double* box, const double* cell, int32_t w, int32_t h, double thr) {
  return pg_edge_touch(box[0], box[1], box[2], box[3], cell[0], cell[1], cell[2], cell[3], (double)w, (double)h, thr) ? 1 : 0;
}
extern "C" double pg_hostcheck_density_weight(int32_t bin, int32_t left, int32_t right, int32_t center) {
  return pg_density_weight(bin, left, right, center);
}
extern "C" int64_t pg_hostcheck_density_rcp_mismatches(int32_t max_span, int32_t max_n) {
  int64_t bad = 0;
  for (int32_t span = 0; span <= max_span; ++span) {
    const double half = pg_density_half(0, span);
    const double inv = 1.0 / half;
    for (int32_t n = 0; n <= max_n; ++n)
      bad += pg_density_weight_rcp(n, 0, half, inv) != pg_density_weight(n, 0, span, 0) ? 1 : 0;
  }
  return bad;
}
// one output row (BGR uint8, dst_w pixels) of the fixed-point resize from the two source rows it needs
extern "C" int pg_hostcheck_resize_row(const uint8_t* row0, const uint8_t* row1, int32_t src_w, int32_t src_h,
                                       int32_t dst_w, int32_t dst_h, int32_t dy, uint8_t* out_bgr) {
  const PgCoef cy = pg_resize_coef(src_h, dst_h, dy, false);
  for (int x = 0; x < dst_w; ++x) {
    const PgCoef cx = pg_resize_coef(src_w, dst_w, x, true);
    for (int c = 0; c < 3; ++c) {
      const uint32_t h0 = pg_hpass(row0[3 * cx.s0 + c], row0[3 * cx.s1 + c], cx.c0, cx.c1);
      const uint32_t h1 = pg_hpass(row1[3 * cx.s0 + c], row1[3 * cx.s1 + c], cx.c0, cx.c1);
      out_bgr[3 * x + c] = (uint8_t)pg_vpass(h0, h1, cy.c0, cy.c1);
    }
  }
  return cy.s0 | (cy.s1 << 16);
}
