/* pagegeom.h — C ABI of libpagegeom.so, the B200 (sm_100a) implementation of the
 * page-geometry hot path of calhounpaul/multimodal_embeddings:
 *
 *     tile -> edge filter -> cross-tile NMS merge -> width median -> column centres
 *
 * The reference is a set of Python scripts with no FFI seam; each entry point
 * below names the reference function(s) (file:line, relative to the upstream repo
 * root) whose arithmetic it replaces.  INTEGRATION.md shows the ctypes stubs a
 * maintainer of the reference would add to call them from the numbered scripts.
 *
 * Conventions
 *   - plain C types only; every pointer marked "dev" is a CUDA device pointer owned
 *     by the caller, everything else is host memory;
 *   - every launch is asynchronous on `stream` (a cudaStream_t passed as void*);
 *   - return value: PG_OK or an error code; pg_last_error() gives the text
 *     (thread-local).  Nothing throws, nothing falls back to the CPU;
 *   - batching: boxes of many pages are concatenated; `page_off[P+1]` (dev, int64)
 *     gives each page's slice; a stage may restrict itself to a subset through
 *     `sel_idx` (dev, int32 global box indices, page p's entries stored from
 *     page_off[p]) and `n_sel[P]` (dev) — exactly what the previous stage emitted —
 *     so the stages chain on-device without host round trips;
 *   - all box arithmetic is IEEE fp64 in the reference's operation order with FMA
 *     contraction disabled (bit-exact against the Python doubles).
 */
#ifndef PAGEGEOM_H_
#define PAGEGEOM_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PG_OK 0
#define PG_ERR_INVALID 1     /* bad argument */
#define PG_ERR_CUDA 2        /* CUDA runtime error (text in pg_last_error) */
#define PG_ERR_WORKSPACE 3   /* caller workspace too small */
#define PG_ERR_UNSUPPORTED 4 /* shape outside what the kernels handle */

#define PG_FLAG_PLAIN_TEXT 1u /* class_name == 'plain_text' (4_extract_median_widths.py:136) */
#define PG_FLAG_TITLE 2u      /* class_name == 'title'      (5_detect_column_centers.py:111) */

#define PG_WIDTH_HIST_BINS 16384 /* corpus width histogram: 1-px bins            */
#define PG_COL_HIST_BINS 1001    /* corpus column histogram: per-mille of page W */
#define PG_COL_SPAN_BYTES 32     /* pg_column_peaks scratch: bytes per input box      */

const char* pg_last_error(void);
int pg_version(void);
/* 1 when the library was built with -DPG_CHECKED (device-side bounds assertions, csrc/pg_common.cuh), else 0. */
int pg_build_checked(void);
/* Page-locked host staging for what crosses PCIe every step (the scans' file bytes; reference: the bytes cv2.imread
 * reads from disk, 1_doclayout_bboxes.py:381).  write_combined != 0: cudaHostAllocWriteCombined (the CPU only fills
 * it; device reads do not snoop the CPU caches).  Free with pg_pinned_free. */
int pg_pinned_alloc(size_t bytes, int32_t write_combined, void** out);
int pg_pinned_free(void* p);
/* SM count / compute capability of the current device. */
int pg_device_info(int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor);

/* ------------------------------------------------------------------ K1 tiler
 * Replaces split_image_into_grid (1_doclayout_bboxes.py:366-444) and the front half
 * of YOLODocumentLayoutDetector.detect_regions (1_doclayout_bboxes.py:191-210:
 * third-party LetterBox -> cv2.resize(INTER_LINEAR) -> pad 114 -> BGR->RGB -> CHW
 * -> /255), without the PNG round trip of 1_doclayout_bboxes.py:568,202.
 */
typedef struct PgTilePlan PgTilePlan;

typedef struct PgTileInfo {
  double x_start, y_start, x_end, y_end; /* un-truncated cell rectangle (:401-421)  */
  int32_t grid_rows, grid_cols;          /* grid this tile belongs to               */
  int32_t row, col;                      /* 1-indexed (:440-441)                    */
  int32_t x0, y0, x1, y1;                /* int()-truncated slice (:424-430)        */
  int32_t new_w, new_h;                  /* resized size inside the letterbox       */
  int32_t pad_l, pad_t;                  /* 114-valued border, left / top           */
  int32_t out_w, out_h;                  /* tile tensor is [3, out_h, out_w] fp16   */
  int64_t out_offset;                    /* in fp16 elements from the page's output */
} PgTileInfo;

/* One plan per page size.  `grid_rows/grid_cols[n_grids]` lists the grids applied to
 * every page (a 1x1 grid is the reference's full-page pass, :446-482).  Tiles are
 * enumerated grid-major, then row-major like the reference.  Host-only; no device
 * work happens until the first pg_tile_letterbox call. */
int pg_tile_plan_create(int32_t page_w, int32_t page_h, const int32_t* grid_rows,
                        const int32_t* grid_cols, int32_t n_grids, double overlap_percentage,
                        int32_t imgsz, int32_t stride, int32_t auto_pad, int32_t scaleup,
                        PgTilePlan** plan);
/* The same for pages of `channels` bytes per pixel: 3 = BGR interleaved (what pg_tile_plan_create builds), 1 = ONE
 * grey plane — a greyscale scan decoded on the device (pg_jpeg_decode), for which cv2.imread would have replicated
 * the plane into three equal channels (1:381).  The one-channel kernel reads the plane once and writes the same
 * value to the three output planes: tiles bit-identical to the three-channel path on the replicated page, a third
 * of the page bytes.  pitch >= channels*W; algorithmic bytes = channels*W*H + 2*out_elems. */
int pg_tile_plan_create_ex(int32_t page_w, int32_t page_h, int32_t channels, const int32_t* grid_rows,
                           const int32_t* grid_cols, int32_t n_grids, double overlap_percentage,
                           int32_t imgsz, int32_t stride, int32_t auto_pad, int32_t scaleup,
                           PgTilePlan** plan);
void pg_tile_plan_destroy(PgTilePlan* plan);
int32_t pg_tile_plan_num_tiles(const PgTilePlan* plan);
int pg_tile_plan_tile(const PgTilePlan* plan, int32_t tile, PgTileInfo* info);
int64_t pg_tile_plan_out_elems(const PgTilePlan* plan);      /* fp16 elements per page  */
int64_t pg_tile_plan_algorithmic_bytes(const PgTilePlan* plan); /* 3WH + 2*out_elems    */

/* pages: dev uint8 BGR, HWC, `n_pages` pages `page_stride` bytes apart, rows `pitch`
 * bytes apart (pitch % 16 == 0, base 16-byte aligned — bulk-copy alignment).
 * out: dev fp16, page p's tiles start at out + p*out_page_stride (elements).
 * Persistent warp-specialised kernel: TMA bulk copies stage the two source rows of
 * each output row in shared memory behind an mbarrier ring. */
int pg_tile_letterbox(PgTilePlan* plan, const uint8_t* pages, int32_t n_pages, int64_t pitch,
                      int64_t page_stride, void* out_f16, int64_t out_page_stride, void* stream);
/* Validation twin: same arithmetic with plain global loads, no staging.  Used by the
 * tests as an independent on-device cross-check at full size. */
int pg_tile_letterbox_direct(PgTilePlan* plan, const uint8_t* pages, int32_t n_pages, int64_t pitch,
                             int64_t page_stride, void* out_f16, int64_t out_page_stride, void* stream);

/* Heterogeneous batches: pages of different sizes (one plan per distinct size) tiled by ONE launch —
 * what a real corpus looks like (the reference's 19 scans have 19 sizes).  page_plan[i] names the plan
 * of page i; bind() records each page's device pointer, pitch and output buffer (pg_tile_plan_out_elems
 * of its plan); the plans may be destroyed after create().  Same kernel, same arithmetic. */
typedef struct PgTileBatch PgTileBatch;
int pg_tile_batch_create(const PgTilePlan* const* plans, int32_t n_plans, const int32_t* page_plan,
                         int32_t n_pages, PgTileBatch** batch);
void pg_tile_batch_destroy(PgTileBatch* batch);
int64_t pg_tile_batch_algorithmic_bytes(const PgTileBatch* batch);
int pg_tile_batch_bind(PgTileBatch* batch, const uint8_t* const* page_ptrs /*host array of dev ptrs*/,
                       const int64_t* pitches /*host [P]*/, void* const* out_ptrs /*host array of dev ptrs*/,
                       void* stream);
int pg_tile_letterbox_batch(PgTileBatch* batch, void* stream);

/* Synthetic newspaper-like pages generated on the device (counter-based hash of
 * (seed0 + first_page + p, y, x)); used by bench.py so inputs are HBM-resident. */
int pg_synth_pages(uint8_t* pages, int32_t n_pages, int32_t page_w, int32_t page_h, int64_t pitch,
                   int64_t page_stride, uint64_t seed0, int64_t first_page, void* stream);

/* ------------------------------------------------------------------ K2 edge filter
 * Replaces translate_coordinates_to_original (1_doclayout_bboxes.py:484-511) and
 * is_box_touching_internal_edge / filter_grid_info (2_edge_box_filter.py:44-90,
 * 206-217).  boxes_are_local != 0: `boxes` are cell-local detector outputs and
 * the float cell origin is added first (result optionally stored to boxes_page_out);
 * else `boxes` are already page coordinates.  kept_idx keeps input order. */
int pg_edge_filter(const double* boxes /*dev [N,4]*/, int32_t boxes_are_local,
                   const int32_t* box_cell /*dev [N] index into cells*/,
                   const double* cells /*dev [C,4] x_start,y_start,x_end,y_end*/,
                   const int32_t* page_wh /*dev [P,2] width,height*/,
                   const int64_t* page_off /*dev [P+1]*/, int32_t n_pages, double threshold,
                   double* boxes_page_out /*dev [N,4] or NULL*/, uint8_t* keep /*dev [N] or NULL*/,
                   int32_t* kept_idx /*dev [N]*/, int32_t* n_kept /*dev [P]*/, void* stream);

/* ------------------------------------------------------------------ K3 NMS merge
 * Replaces calculate_iou / apply_non_max_suppression (3_combine_grids.py:46-138) on
 * the pooled boxes of combine_boxes_for_image (3_combine_grids.py:222-267).
 * Output: kept_idx = global box indices in pick order (score descending, earlier
 * pooled position first on ties), page p's picks stored from page_off[p].
 * max_boxes_per_page is a scheduling hint (0 = unknown), never a correctness input: from 32768 boxes
 * on a page, and while n_pages * 8 fits the SMs, the per-page kernels (bin, resolve, emit) run on one
 * thread-block cluster per page instead of one CTA per page; results are identical either way.
 * Environment knobs read at each call: PG_NMS_CLUSTER_MIN_BOXES (that threshold; <= 0 disables the
 * cluster kernels), PG_NMS_MASK_OCC (3 or 4 resident CTAs per SM for the mask kernel). */
size_t pg_nms_workspace_bytes(int64_t n_boxes, int32_t n_pages, int32_t pairs_per_block);
int pg_nms_merge(const double* boxes /*dev [N,4]*/, const double* scores /*dev [N]*/,
                 const double* classes /*dev [N]*/, const int32_t* sel_idx /*dev or NULL*/,
                 const int64_t* page_off /*dev [P+1]*/, const int32_t* n_sel /*dev [P] or NULL*/,
                 int32_t n_pages, int64_t n_boxes, int32_t max_boxes_per_page, double iou_threshold,
                 int32_t* kept_idx /*dev [N]*/, int32_t* n_kept /*dev [P]*/,
                 void* workspace /*dev*/, size_t workspace_bytes, void* stream);
/* Same machinery with mode bits.  PG_NMS_CLASS_AGNOSTIC | PG_NMS_FP32 reproduces
 * torchvision.ops.nms as the reference calls it per tile (1_doclayout_bboxes.py:217-225): boxes and
 * scores are float32 values (passed here widened to double), areas / intersection / IoU are evaluated
 * in float32, ties keep the earlier box, classes are ignored (may be NULL). */
#define PG_NMS_CLASS_AGNOSTIC 1
#define PG_NMS_FP32 2
int pg_nms_merge_ex(const double* boxes, const double* scores, const double* classes,
                    const int32_t* sel_idx, const int64_t* page_off, const int32_t* n_sel,
                    int32_t n_pages, int64_t n_boxes, int32_t max_boxes_per_page, double iou_threshold,
                    int32_t mode, int32_t* kept_idx, int32_t* n_kept, void* workspace,
                    size_t workspace_bytes, void* stream);
/* After the stream has been synchronised: status word + counters the merge left in
 * its workspace. stats[0]=status (PG_OK/PG_ERR_*), [1]=candidate block pairs,
 * [2]=resolve rounds (max over pages), [3]=box pairs that went through the prefilter test (counted in the
 * kernel: runs of J skipped by their bounds are not in it). */
int pg_nms_stats(const void* workspace /*dev*/, int64_t stats[4]);

/* ------------------------------------------------------------------ K4 width median
 * Replaces bin_widths / calculate_median_width and the plain_text width extraction
 * (4_extract_median_widths.py:49-101, 135-141).  flags: dev [N] PG_FLAG_* per box. */
int pg_class_flags(const double* classes /*dev [N]*/, int64_t n, double plain_text_id,
                   double title_id, uint8_t* flags /*dev [N]*/, void* stream);
int pg_width_median(const double* boxes, const uint8_t* flags, const int32_t* sel_idx,
                    const int64_t* page_off, const int32_t* n_sel, int32_t n_pages, int64_t n_boxes,
                    const int32_t* page_wh, double min_margin_percent,
                    double* median /*dev [P]*/, int32_t* n_bins /*dev [P]*/,
                    double* ws_keys /*dev [2N]: bin keys + gathered widths*/, int32_t* ws_counts /*dev [N]*/,
                    uint32_t* width_hist /*dev [PG_WIDTH_HIST_BINS] or NULL*/, void* stream);

/* ------------------------------------------------------------------ K5 column centres
 * Replaces find_column_centers (5_detect_column_centers.py:91-224) including the
 * scipy.signal.find_peaks(height, distance, prominence) step order and np.convolve
 * ('same').  gauss_table holds, for every odd window length M = 2k+1 <= max_window,
 * the normalised scipy.signal.windows.gaussian(M, std=M/6) starting at
 * gauss_off[k] (host-built so that np.exp rounding is shared with the reference).
 * Outputs per page: up to max_cols centres (px, = peak*resolution) and widths. */
int pg_column_peaks(const double* boxes, const uint8_t* flags, const double* scores,
                    const int32_t* sel_idx, const int64_t* page_off, const int32_t* n_sel,
                    int32_t n_pages, const int32_t* page_wh, const double* median /*dev [P]*/,
                    const double* gauss_table /*dev*/, const int64_t* gauss_off /*dev*/,
                    int32_t max_window, double min_confidence, int32_t max_cols,
                    int32_t* centers /*dev [P,max_cols]*/, double* widths /*dev [P,max_cols]*/,
                    int32_t* n_cols /*dev [P]; <0 = unsupported shape*/,
                    double* ws /*dev [P, 2*max_bins]: density, smoothed density*/, int32_t max_bins,
                    void* ws_spans /*dev, PG_COL_SPAN_BYTES * N, 16-byte aligned*/,
                    int32_t* ws_span_counts /*dev [P]*/,
                    uint32_t* col_hist /*dev [PG_COL_HIST_BINS] or NULL*/, void* stream);

/* Per-box column assignment.  The reference stops at centres/widths (SURVEY.md 8a note); this is the
 * documented derived function: col_of_box[i] = index of the centre nearest to (x0+x1)/2 of box i
 * (first minimum on ties), -1 for boxes outside the selection or pages with no column. */
int pg_assign_columns(const double* boxes /*dev [N,4]*/, const int32_t* sel_idx, const int64_t* page_off,
                      const int32_t* n_sel, int32_t n_pages, int64_t n_boxes,
                      const int32_t* centers /*dev [P,max_cols]*/, const int32_t* n_cols /*dev [P]*/,
                      int32_t max_cols, int32_t* col_of_box /*dev [N]*/, void* stream);

/* ------------------------------------------------------------------ K6 corpus-histogram exchange
 * The one exchange step of the path (the reference is a single process and has no analogue):
 * the per-rank integer histograms pg_width_median / pg_column_peaks accumulate (width_hist, col_hist) are
 * summed over the ranks, in place, with ONE ncclAllReduce(ncclUint32, ncclSum) over NVLink.  Integer sums are
 * order-independent: the totals are bit-identical for any number of GPUs.
 *   pg_hist_allreduce(hist, n, comm, stream)   `comm` is an ncclComm_t (passed as void* so that this header does
 *                                              not need nccl.h); asynchronous on `stream`.
 * For callers that do not hold an NCCL communicator: rank 0 makes an id with pg_comm_unique_id, sends the 128
 * bytes to every rank by whatever means the host has (the Python side uses torch.distributed), and every rank
 * calls pg_comm_create (collective); pg_comm_nccl() is the ncclComm_t to pass above.  NCCL is bound at run time
 * (libnccl.so.2, sharing the host process's copy when it has one); without it these return PG_ERR_UNSUPPORTED. */
typedef struct PgComm PgComm;
int pg_comm_nccl_version(void); /* 0: NCCL not available */
int pg_comm_unique_id(uint8_t id[128]);
int pg_comm_create(const uint8_t id[128], int32_t world, int32_t rank, PgComm** comm);
void pg_comm_destroy(PgComm* comm);
void* pg_comm_nccl(PgComm* comm);
int pg_hist_allreduce(uint32_t* hist /*dev [n_bins]*/, size_t n_bins, void* nccl_comm /*ncclComm_t*/, void* stream);

/* ------------------------------------------------------------------ J1-J4 stage-3 record writer
 * Replaces `json.dump(result, f, indent=2)` (3_combine_grids.py:441-443) of the dict built by
 * combine_boxes_for_image (3:282-291): the documents of n_pages pages are laid out on the device, byte for
 * byte as CPython prints them (floats as float.__repr__: shortest round-trip digits, csrc/pg_fmt.h), from the
 * merge's own outputs — kept_idx/n_kept as everywhere (kept_idx NULL: every box of the page).
 *
 * The caller supplies the page-invariant text, already JSON-encoded, in one device blob `text`:
 *   head  text[head_off[p] .. head_off[p+1])  from "{" up to and including  "boxes": [
 *   tail  text[tail_off[p] .. tail_off[p+1])  from the newline before  "source_jsons"  up to the final "}"
 *   name  text[name_off[i] .. name_off[i+1])  class-name string literal i (with its quotes); name_id[box] = i
 * Documents are packed back to back into `out`: page p is out[out_off[p] .. out_off[p+1]).  out_off (device,
 * n_pages+1) is always written; if out_off[n_pages] > out_capacity nothing is written to `out` and the
 * caller retries with a larger buffer.  Workspace: pg_json_workspace_bytes(n_boxes, n_pages), 256-aligned. */
#define PG_JSON_SLOT_BYTES 32
int64_t pg_json_workspace_bytes(int64_t n_boxes, int32_t n_pages);
int pg_json_combined(const double* boxes /*dev [N,4]*/, const double* classes /*dev [N]*/,
                     const double* scores /*dev [N]*/, const int32_t* name_id /*dev [N]*/,
                     const int32_t* kept_idx, const int64_t* page_off, const int32_t* n_kept,
                     int32_t n_pages, int64_t n_boxes, int32_t max_boxes_per_page,
                     const uint8_t* text /*dev*/, const int64_t* head_off /*dev [P+1]*/,
                     const int64_t* tail_off /*dev [P+1]*/, const int64_t* name_off /*dev [n_names+1]*/,
                     uint8_t* out /*dev*/, int64_t out_capacity, int64_t* out_off /*dev [P+1]*/,
                     void* ws, int64_t ws_bytes, void* stream);

/* Generic form of the writer, for the other json.dump(indent=2) schemas of the reference (the nested
 * cells[].regions documents of stages 1/2: 1_doclayout_bboxes.py:592-594,645-647, 2_edge_box_filter.py:485-487).
 * The output is a flat list of segments: a piece of host-encoded text followed by one array printed from device
 * data (nothing after the text for PG_JSON_KIND_TEXT).  A document is any run of consecutive segments; the
 * caller cuts it out of `out` with seg_out_off (dev [S+1], always written; if seg_out_off[S] > out_capacity
 * nothing is written to `out`).  Element k of a segment is box kept_idx[start + k] (or start + k when kept_idx is
 * NULL) of data[data_id]: f64 [N,4] for BOX4, f64 [N] for SCALAR, int32 [N] name ids for NAME (name i is
 * text[name_off[i] .. name_off[i+1]), a JSON string literal).  elem_off (dev [S+1]) = exclusive prefix of the
 * segments' counts (0 for TEXT).  `indent` = spaces in front of an entry (the array's key sits at indent - 2). */
#define PG_JSON_KIND_TEXT 0
#define PG_JSON_KIND_BOX4 1
#define PG_JSON_KIND_SCALAR 2
#define PG_JSON_KIND_NAME 3
typedef struct PgJsonSegment {
  int64_t head_begin, head_end; /* text[head_begin .. head_end) precedes the array */
  int64_t start;                /* first element */
  int32_t count;                /* elements (0: prints "[]") */
  int32_t kind;                 /* PG_JSON_KIND_* */
  int32_t data_id;              /* index into data[] */
  int32_t indent;
} PgJsonSegment;
int64_t pg_json_segments_workspace_bytes(int64_t n_elems, int32_t n_segs);
int pg_json_segments(const PgJsonSegment* segs /*dev [S]*/, int32_t n_segs, const int64_t* elem_off /*dev [S+1]*/,
                     int64_t n_elems, const void* const* data /*host array of <= 8 dev pointers*/, int32_t n_data,
                     const int32_t* kept_idx /*dev or NULL*/, const uint8_t* text /*dev*/,
                     const int64_t* name_off /*dev or NULL when no NAME segment*/, uint8_t* out /*dev*/,
                     int64_t out_capacity, int64_t* seg_out_off /*dev [S+1]*/, void* ws, int64_t ws_bytes,
                     void* stream);

/* ------------------------------------------------------------------ R1-R3 record reader
 * The numbers of a record's "boxes" / "classes" / "scores" arrays, text -> f64 on the device: what json.load's
 * float() does for the reference's stage-4/5 readers (4_extract_median_widths.py:103-151,
 * 5_detect_column_centers.py:337-400), correctly rounded for every token of <= 17 significant digits
 * (everything json.dump prints; csrc/pg_fmt.h).  `ranges` (dev, [R,2] begin/end byte offsets into `text`) are
 * regions holding only numbers, brackets, commas and whitespace — the caller locates them and parses the
 * record's few strings itself.  range_block_off (dev [R+1]) = exclusive prefix of ceil(len_r / PG_JSON_PARSE_BLOCK).
 * Range r's numbers land, in text order, in values[val_off[r] .. val_off[r+1]) (nothing is stored beyond
 * `capacity`; val_off[R] is the total, so a short buffer is detected and the call repeated); n_bad[r] counts
 * tokens that need the host's float() (more than 17 digits, malformed) and, with PG_JSON_INT_LITERALS_TO_HOST,
 * integer literals — json.load makes Python ints of those, and a caller that must re-emit them unchanged
 * ("5", not "5.0") lets CPython read that file. */
#define PG_JSON_INT_LITERALS_TO_HOST 1
#define PG_JSON_PARSE_BLOCK 2048 /* text bytes per CTA (256 threads x 8 consecutive bytes) */
int32_t pg_json_parse_block_bytes(void);
int64_t pg_json_parse_workspace_bytes(int64_t total_blocks);
int pg_json_parse_numbers(const uint8_t* text /*dev*/, const int64_t* ranges /*dev [R,2]*/, int32_t n_ranges,
                          const int64_t* range_block_off /*dev [R+1]*/, int64_t total_blocks,
                          double* values /*dev [capacity]*/, int64_t capacity, int64_t* val_off /*dev [R+1]*/,
                          int32_t* n_bad /*dev [R]*/, int32_t flags, void* ws, int64_t ws_bytes, void* stream);

/* ------------------------------------------------------------------ D1-D8 JPEG scans decoded on the device
 * Replaces, for `.jpg` input, the `cv2.imread(image_path)` that opens the path (1_doclayout_bboxes.py:381, once
 * per grid; 2_edge_box_filter.py:195): the files cross PCIe compressed and the pages are produced in HBM, bit for
 * bit what cv2 (libjpeg-turbo: Huffman decode, dequantisation, "islow" integer IDCT, fancy chroma upsampling,
 * fixed-point YCbCr -> RGB) returns.  Baseline / extended sequential Huffman, 8-bit, one scan; greyscale files give
 * ONE grey plane (cv2 replicates it into three equal channels; the tiler's one-channel plans read the single plane
 * instead), YCbCr files with 4:4:4 / 4:2:2 / 4:2:0 / 4:4:0 sampling give BGR interleaved rows like cv2.  Anything
 * else — progressive, arithmetic coding, 12-bit, CMYK — is PG_ERR_UNSUPPORTED at set_files time and the caller keeps
 * its host decoder for that file.
 *   set_files   host: parses the headers of n files lying back to back in `blob` (host memory; file i is
 *               blob[file_off[i] .. file_off[i+1])); nothing is copied, the device gets the same blob.
 *   decode      device, asynchronous: blob_dev = the same bytes in device memory; page i is written to
 *               out_ptrs[i] as uint8 [height, pitches[i]] (16-byte aligned, pitch % 16 == 0, pitch >= channels * width).
 *               blob_dev must be 16-byte aligned and readable up to the next multiple of 16 behind its last file.
 *   status      after the stream has been synchronised: stats[0] = PG_OK, or PG_ERR_UNSUPPORTED when the
 *               configured number of sync rounds (default 3; pg_jpeg_decoder_configure) did not reach the fixed
 *               point of the chunk states — raise it and decode again; [1] = rounds that changed a state,
 *               [2] = states replaced in round 1, [3] = chunks.
 * Entropy decoding is parallel over fixed chunks of the unstuffed stream (default 512 bytes), self-synchronising;
 * csrc/pg_jpeg.h describes the passes. */
typedef struct PgJpegDecoder PgJpegDecoder;
int pg_jpeg_decoder_create(PgJpegDecoder** dec);
void pg_jpeg_decoder_destroy(PgJpegDecoder* dec);
int pg_jpeg_decoder_configure(PgJpegDecoder* dec, int32_t chunk_bytes, int32_t sync_rounds);
int pg_jpeg_decoder_set_files(PgJpegDecoder* dec, const uint8_t* blob /*host*/, const int64_t* file_off /*host [n+1]*/,
                              int32_t n);
int pg_jpeg_decoder_image_info(const PgJpegDecoder* dec, int32_t i, int32_t* width, int32_t* height, int32_t* channels);
int64_t pg_jpeg_workspace_bytes(const PgJpegDecoder* dec); /* for the files of the last set_files */
int pg_jpeg_decode(PgJpegDecoder* dec, const uint8_t* blob_dev, uint8_t* const* out_ptrs /*host array of dev ptrs*/,
                   const int64_t* pitches /*host [n]*/, void* workspace /*dev, 256-aligned*/, int64_t workspace_bytes,
                   void* stream);
/* Optional, for callers that overlap the upload of the next batch's files with the kernels of this one: uploads the
 * batch's tables (a few hundred KB) into `workspace` on `copy_stream` — the stream that carries the files — so that
 * the decode issues no host->device copy of its own.  Host->device copies share one DMA engine: a small table copy
 * issued with the kernels would queue behind the other batch's 300 MB of files and stall the kernels for that long
 * (measured: 21.5 instead of 11.3 ms per 32-page step).  The pg_jpeg_decode that follows takes the same outputs and
 * workspace and must be ordered after this call (an event on copy_stream). */
int pg_jpeg_stage_tables(PgJpegDecoder* dec, uint8_t* const* out_ptrs, const int64_t* pitches, void* workspace,
                         int64_t workspace_bytes, void* copy_stream);
int pg_jpeg_decode_status(const PgJpegDecoder* dec, int64_t stats[4]);

/* ------------------------------------------------------------------ test hooks
 * Host evaluations of the same inline arithmetic the kernels are compiled from
 * (csrc/pg_math.h, csrc/pg_fmt.h).  Used by the CPU test-suite only; not a compute path. */
/* float.__repr__ of x into buf (>= 24 bytes, no terminator); returns the length */
int32_t pg_hostcheck_format_double(double x, char* buf);
/* the same for n values, value i at buf + 24*i with length len[i]; returns the total length */
int64_t pg_hostcheck_format_doubles(const double* x, int64_t n, char* buf, int32_t* len);
/* the reader's number conversion (csrc/pg_fmt.h: pg_parse_json_number): token i starts at text + tok_off[i];
 * out[i] = its value, consumed[i] = bytes used (0: not convertible here, host fallback); returns the count > 0 */
int64_t pg_hostcheck_parse_numbers(const char* text, int64_t text_len, const int64_t* tok_off, int64_t n,
                                   double* out, int32_t* consumed);
double pg_hostcheck_iou(const double* a, const double* b);
/* the divide-free predicate the merge kernel uses for `iou > thr` (must equal pg_hostcheck_iou(a,b) > thr) */
int32_t pg_hostcheck_iou_gt(const double* a, const double* b, double thr);
/* the float32 predicate of the per-tile NMS mode (a suppresses b?) */
int32_t pg_hostcheck_iou_gt_f32(const float* a, const float* b, double thr);
int32_t pg_hostcheck_edge_touch(const double* box, const double* cell, int32_t w, int32_t h, double thr);
double pg_hostcheck_density_weight(int32_t bin, int32_t left, int32_t right, int32_t center);
/* number of (right-left, |bin-center|) pairs in [0,max_span]x[0,max_n] where the reciprocal+FMA
 * form used by the density kernel differs from the true divide (must be 0) */
int64_t pg_hostcheck_density_rcp_mismatches(int32_t max_span, int32_t max_n);
/* the decoder's inline code (csrc/pg_jpeg.h) evaluated on the host, chunk by chunk in the kernels' order: one
 * greyscale file -> out[height, pitch].  stats: [0] sync rounds that changed a state, [1] states replaced in
 * round 1, [2] chunks, [3] restart markers.  out NULL: header only (width, height; stats[3] = components). */
int pg_hostcheck_jpeg_decode(const uint8_t* file, int64_t len, int32_t chunk_bytes, int32_t max_rounds, uint8_t* out,
                             int64_t pitch, int32_t* width, int32_t* height, int64_t stats[4]);
int pg_hostcheck_resize_row(const uint8_t* row0, const uint8_t* row1, int32_t src_w, int32_t src_h,
                            int32_t dst_w, int32_t dst_h, int32_t dy, uint8_t* out_bgr);

#ifdef __cplusplus
}
#endif
#endif /* PAGEGEOM_H_ */
