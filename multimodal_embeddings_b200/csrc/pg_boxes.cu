// pg_boxes.cu — K2 edge filter, K3 cross-tile NMS merge, K4 width median, K5 column centres.
//
// All four are fp64, bit-exact restatements of the reference's Python-double arithmetic
// (pg_math.h), batched over pages through CSR offsets so that one launch serves a whole
// shard.  None of them is HBM-bound (a page's boxes are a few hundred KB); they are organised
// so that the dependent chain per page is short and pages run side by side on different SMs.
#include <cfloat>
#include <cstdlib>

#include <cooperative_groups.h>

#include "pg_common.cuh"

namespace cg = cooperative_groups;

// =============================================================================================
// K2 — edge-touch filter (+ fused cell->page translation)
//   reference: translate_coordinates_to_original 1_doclayout_bboxes.py:484-511,
//              is_box_touching_internal_edge 2_edge_box_filter.py:44-90,
//              filter_grid_info 2_edge_box_filter.py:206-217 (kept order = input order)
// One CTA per page; order-preserving compaction by block scan.
// =============================================================================================
constexpr int EDGE_THREADS = 1024;
// Threads of the one-CTA-per-page kernels whose code does not fix the width (edge filter, resolve, emit).  A CTA
// of 1024 threads needs most of an SM's register file, so under the tiler (four resident CTAs fill it) the SM has to
// drain before such a CTA starts; narrower CTAs slot into what one retiring tiler CTA frees.  Knob: PG_BOX_PAGE_THREADS.
static int page_kernel_threads(int dflt) {
  int t = dflt;
  if (const char* e = getenv("PG_BOX_PAGE_THREADS")) t = atoi(e);
  t = (t / 32) * 32;
  return t < 128 ? 128 : (t > 1024 ? 1024 : t);
}
__global__ void __launch_bounds__(EDGE_THREADS) edge_filter_kernel(
    const double* __restrict__ boxes, int local, const int32_t* __restrict__ box_cell,
    const double* __restrict__ cells, const int32_t* __restrict__ page_wh, const int64_t* __restrict__ page_off,
    double thr, double* __restrict__ boxes_out, uint8_t* __restrict__ keep, int32_t* __restrict__ kept_idx,
    int32_t* __restrict__ n_kept) {
  __shared__ int scan_smem[34];
  const int p = blockIdx.x;
  const int64_t b0 = page_off[p], b1 = page_off[p + 1];
  const double W = (double)page_wh[2 * p], H = (double)page_wh[2 * p + 1];
  int running = 0;
  for (int64_t base = b0; base < b1; base += blockDim.x) {
    const int64_t i = base + threadIdx.x;
    int k = 0;
    if (i < b1) {
      const double2 lo = *reinterpret_cast<const double2*>(boxes + 4 * i);
      const double2 hi = *reinterpret_cast<const double2*>(boxes + 4 * i + 2);
      const int c = box_cell[i];
      const double2 c0 = *reinterpret_cast<const double2*>(cells + 4 * (int64_t)c);
      const double2 c1 = *reinterpret_cast<const double2*>(cells + 4 * (int64_t)c + 2);
      double x0 = lo.x, y0 = lo.y, x1 = hi.x, y1 = hi.y;
      if (local) {  // 1_doclayout_bboxes.py:503-506: box + float cell origin
        x0 = x0 + c0.x; y0 = y0 + c0.y; x1 = x1 + c0.x; y1 = y1 + c0.y;
      }
      if (boxes_out) {
        *reinterpret_cast<double2*>(boxes_out + 4 * i) = make_double2(x0, y0);
        *reinterpret_cast<double2*>(boxes_out + 4 * i + 2) = make_double2(x1, y1);
      }
      k = pg_edge_touch(x0, y0, x1, y1, c0.x, c0.y, c1.x, c1.y, W, H, thr) ? 0 : 1;
      if (keep) keep[i] = (uint8_t)k;
    }
    int total;
    const int ex = pg_block_exscan(k, scan_smem, &total);
    if (k) kept_idx[b0 + running + ex] = (int32_t)i;
    running += total;
  }
  if (threadIdx.x == 0) n_kept[p] = running;
}

extern "C" int pg_edge_filter(const double* boxes, int32_t boxes_are_local, const int32_t* box_cell,
                              const double* cells, const int32_t* page_wh, const int64_t* page_off,
                              int32_t n_pages, double threshold, double* boxes_page_out, uint8_t* keep,
                              int32_t* kept_idx, int32_t* n_kept, void* stream) {
  PG_REQUIRE(n_pages >= 0, "n_pages");
  if (n_pages == 0) return PG_OK;
  PG_REQUIRE(boxes && box_cell && cells && page_wh && page_off && kept_idx && n_kept, "null device pointer");
  edge_filter_kernel<<<n_pages, page_kernel_threads(EDGE_THREADS), 0, (cudaStream_t)stream>>>(
      boxes, boxes_are_local, box_cell, cells, page_wh, page_off, threshold, boxes_page_out, keep, kept_idx, n_kept);
  PG_LAUNCH_CHECK();
  return PG_OK;
}

// =============================================================================================
// K3 — class-aware greedy NMS, exact
//   reference: calculate_iou 3_combine_grids.py:46-78, apply_non_max_suppression :80-138.
//
// Greedy NMS keeps box i iff no *kept* box j with higher priority (score desc, earlier pooled
// position on ties, :112), same class and IoU > thr (:130) exists.  That fixed point is unique,
// so it can be computed without replaying the sequential loop:
//   A  bin      per page, a stable block radix sort (two counting-sort passes; three on crowded pages, with a
//               finer y digit first) orders the boxes by (x strip, y cell[, y sub-cell]) of their centres so that
//               runs of 32 consecutive boxes ("blocks") are spatially compact; write a blocked AoS copy, the block
//               bounding boxes, the bounds of each run of 32 blocks ("super-blocks"), and reset the call's counters;
//   B  cand     per block I, the blocks J of the same page whose bounding boxes intersect (IoU > thr >= 0 needs
//               a non-empty intersection, the reference's own early-out), found through the super-blocks; one
//               atomicAdd reserves the block's entry range (workspace overflow -> status), the hits of the
//               counting walk are kept in shared memory and copied;
//   C  mask     per candidate pair (I,J): lane i of the warp holds box i of I, the 32 boxes of J
//               are broadcast from shared memory; lane i accumulates a 32-bit mask of the boxes
//               of J that would suppress it;
//   D  resolve  per page, Jacobi rounds over two bit-words per block (kept / undecided): an
//               undecided box becomes suppressed if a suppressor is kept, kept if none of its
//               suppressors is still undecided.  Each round decides at least the best undecided
//               box, typical depth is 3-8 rounds; blocks that are final leave the rounds;
//   E  emit     sort the kept boxes by priority (normalised bitonic network, eight elements to a thread in shared
//               memory, on the L2-resident arrays above 8192 survivors of a one-CTA page) and write global
//               indices in pick order.
// Five launches: nms_bin, nms_cand, nms_mask, nms_resolve, nms_emit.
// Pages of >= 32 768 boxes (cfg4: 100 000) would leave A, D and E on one CTA per page; when the pages
// of a launch number fewer than the SMs, they instead run on one thread-block CLUSTER per page
// (up to 8 CTAs — 16 where every page's cluster stays resident —, cluster barrier between sort passes / rounds /
// network stages, totals, flags and counts exchanged through distributed shared memory):
// nms_bin_cluster_kernel, nms_resolve_cluster_kernel, nms_emit_cluster_kernel.
// =============================================================================================
constexpr int NMS_GY = 128;     // y cells (7-bit digit)
constexpr int NMS_GYF = 128;    // subdivisions of a y cell (a third, finer 7-bit digit), used on crowded pages only
constexpr int NMS_FINE_MIN_PER_CELL = 16;  // ... those with more boxes per (x strip, y cell) than this on average
constexpr int NMS_GX_MAX = 64;  // x strips (6-bit digit), chosen per page from the mean box width
constexpr int NMS_CLUSTER_MAX = 16;  // CTAs per page when a launch has few, large pages (above 8: non-portable cluster size)

struct __align__(16) SBox {  // one box in spatial order
  double x0, y0, x1, y1;
  double area, score, cls;
  long long k;  // position in the pooled list (priority tie-break), -1 = padding lane
};

struct NmsWs {
  int64_t* stats;      // [8]: status, candidate pairs, rounds, box pairs tested
  int32_t* sorted_pos; // [N]   spatial order -> local position k
  int32_t* cellid;     // [N]
  SBox* sbox;          // [NB*32]
  double* bbox;        // [NB*4]
  double* sbbox;       // [(NB/32 + P + 1)*4]  bounds of each run of 32 blocks of a page ("super-block")
  float4* bboxf;       // [NB]    the same, rounded outwards to fp32 (conservative prefilter only)
  float4* subbox;      // [NB*4]  outward-rounded fp32 bounds of each run of 8 boxes inside a block
  int32_t* blk_page;   // [NB]
  int32_t* cand_cnt;   // [NB]
  int64_t* cand_off;   // [NB+1]
  uint32_t* st_kept;   // [2*NB]
  uint32_t* st_undec;  // [2*NB]
  double* kscore;      // [N]  (re-used as sortable u64 keys by emit)
  int32_t* kpos;       // [N]  (scratch of the radix sort, then kept positions)
  int32_t* ent_j;      // [E]  candidate block J (page-relative), -1 once its masks are all zero
  int32_t* ent_i;      // [E]  block I the entry belongs to (global block id)
  uint32_t* ent_mask;  // [E*32]
  int64_t nb_cap, ent_cap;
  int32_t mode;        // PG_NMS_* bits
  int32_t fine_min;    // boxes per (x strip, y cell) above which the sub-cell digit is sorted too
};
// stats[] slots
enum { ST_STATUS = 0, ST_PAIRS = 1, ST_ROUNDS = 2, ST_TESTS = 3, ST_ENT_TOTAL = 4 };

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static size_t nms_layout(int64_t n, int32_t n_pages, int32_t pairs_per_block, uint8_t* base, NmsWs* ws) {
  const int64_t nb = (n >> 5) + n_pages + 1;
  const int64_t ecap = (int64_t)(pairs_per_block > 0 ? pairs_per_block : 64) * nb;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    uint8_t* p = base ? base + off : nullptr;
    off = align_up(off + bytes, 256);
    return p;
  };
  NmsWs w;
  w.stats = (int64_t*)take(8 * sizeof(int64_t));
  w.sorted_pos = (int32_t*)take((size_t)n * 4);
  w.cellid = (int32_t*)take((size_t)n * 4);
  w.sbox = (SBox*)take((size_t)nb * 32 * sizeof(SBox));
  w.bbox = (double*)take((size_t)nb * 4 * 8);
  w.sbbox = (double*)take((size_t)((nb >> 5) + n_pages + 1) * 4 * 8);
  w.bboxf = (float4*)take((size_t)nb * sizeof(float4));
  w.subbox = (float4*)take((size_t)nb * 4 * sizeof(float4));
  w.blk_page = (int32_t*)take((size_t)nb * 4);
  w.cand_cnt = (int32_t*)take((size_t)nb * 4);
  w.cand_off = (int64_t*)take((size_t)(nb + 1) * 8);
  w.st_kept = (uint32_t*)take((size_t)nb * 2 * 4);
  w.st_undec = (uint32_t*)take((size_t)nb * 2 * 4);
  w.kscore = (double*)take((size_t)n * 8);
  w.kpos = (int32_t*)take((size_t)n * 4);
  w.ent_j = (int32_t*)take((size_t)ecap * 4);
  w.ent_i = (int32_t*)take((size_t)ecap * 4);
  w.ent_mask = (uint32_t*)take((size_t)ecap * 32 * 4);
  w.nb_cap = nb;
  w.ent_cap = ecap;
  if (ws) *ws = w;
  return off;
}

extern "C" size_t pg_nms_workspace_bytes(int64_t n_boxes, int32_t n_pages, int32_t pairs_per_block) {
  if (n_boxes < 0 || n_pages < 0) return 0;
  return nms_layout(n_boxes, n_pages, pairs_per_block, nullptr, nullptr);
}

struct PageSpan {
  int64_t base;  // first slot of this page in the pooled arrays
  int32_t m;     // boxes selected on this page
  int64_t blk0;  // first block id
  int32_t nb;    // blocks used
};

__device__ __forceinline__ PageSpan page_span(const int64_t* page_off, const int32_t* n_sel, int p) {
  PageSpan s;
  s.base = page_off[p];
  s.m = n_sel ? n_sel[p] : (int32_t)(page_off[p + 1] - s.base);
  s.blk0 = (s.base >> 5) + p;
  s.nb = (s.m + 31) >> 5;
  return s;
}

__device__ __forceinline__ double warp_min_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_max_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// One stable counting-sort pass of a 1024-thread CTA (32 warps, warp w owns a contiguous chunk
// of the input so that (digit, warp, position) order == (digit, input order)).
//   hist: shared int[32 * NDIG]; elem(i) -> element id of input slot i; digit(e) -> [0, NDIG)
template <int NDIG, typename ElemOf, typename DigitOf>
__device__ __forceinline__ void block_stable_pass(int m, int* hist, int* scan_smem, ElemOf elem, DigitOf digit,
                                                  int32_t* dst) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < 32 * NDIG; i += 1024) hist[i] = 0;
  __syncthreads();
  const int chunk = (((m + 31) >> 5) + 31) & ~31;
  const int lo = min(m, warp * chunk), hi = min(m, lo + chunk);
  for (int i = lo + lane; i < hi; i += 32) atomicAdd(&hist[warp * NDIG + digit(elem(i))], 1);
  __syncthreads();
  {
    constexpr int PER = (32 * NDIG) / 1024;
    int v[PER], sum = 0;
#pragma unroll
    for (int q = 0; q < PER; ++q) {
      const int idx = tid * PER + q;  // (digit major, warp minor)
      v[q] = hist[(idx & 31) * NDIG + (idx >> 5)];
      sum += v[q];
    }
    int total;
    int ex = pg_block_exscan(sum, scan_smem, &total);
#pragma unroll
    for (int q = 0; q < PER; ++q) {
      const int idx = tid * PER + q;
      hist[(idx & 31) * NDIG + (idx >> 5)] = ex;
      ex += v[q];
    }
  }
  __syncthreads();
  for (int c = lo; c < hi; c += 32) {
    const int i = c + lane;
    const bool act = i < hi;
    const int e = act ? elem(i) : 0;
    const int d = act ? digit(e) : -1 - lane;
    const unsigned grp = __match_any_sync(0xffffffffu, d);
    const int leader = __ffs(grp) - 1;
    const int rank = __popc(grp & ((1u << lane) - 1u));
    int basepos = 0;
    if (act && lane == leader) {
      basepos = hist[warp * NDIG + d];
      hist[warp * NDIG + d] = basepos + __popc(grp);
    }
    basepos = __shfl_sync(0xffffffffu, basepos, leader);
    PG_DEV_ASSERT(!act || (basepos + rank >= 0 && basepos + rank < m && e >= 0 && e < m));
    if (act) dst[basepos + rank] = e;
    __syncwarp();
  }
  __syncthreads();
}

// Spatial key of a box centre at grid position (fx, fy): x strip | y cell | y sub-cell, 6 + 7 + 7 bits.  Pure
// heuristic — any ordering gives the same kept set; a tighter one gives fewer candidate block pairs.  The sub-cell
// digit costs a third sorting pass and is used only where a (strip, cell) holds many boxes: there, 32 consecutive
// boxes in (strip, cell) order span the whole cell and every block of a cell is a candidate of every other.
__device__ __forceinline__ int32_t nms_cell_id(double fx, double fy, int gxn) {
  const int gx = (fx >= 0.0) ? (fx < (double)(gxn - 1) ? (int)fx : gxn - 1) : 0;
  const double fyc = (fy >= 0.0) ? fy : 0.0;
  const int gy = fyc < (double)(NMS_GY - 1) ? (int)fyc : NMS_GY - 1;
  const double sub = (fyc - (double)gy) * (double)NMS_GYF;
  const int gyf = sub < (double)(NMS_GYF - 1) ? (int)sub : NMS_GYF - 1;
  return (gx * NMS_GY + gy) * NMS_GYF + gyf;
}
__device__ __forceinline__ bool nms_fine_cells(int m, int gxn, int min_per_cell) {
  return (int64_t)m > (int64_t)min_per_cell * NMS_GY * gxn;
}

// ---- A: bin --------------------------------------------------------------------------------
__device__ __forceinline__ void nms_blocked_copy(const double* __restrict__ boxes, const double* __restrict__ scores,
                                                 const double* __restrict__ classes, const int32_t* __restrict__ sel_idx,
                                                 const PageSpan& sp, const int32_t* sorted, const NmsWs& ws, int p,
                                                 int wfirst, int wstride);

// Bounds of each run of 32 blocks of page p ("super-block"; page p's super-blocks start at (blk0 >> 5) + p), by the
// warps wfirst, wfirst + wstride, ... once every block bound of the page is written and visible.
__device__ __forceinline__ void nms_super_bounds(const PageSpan& sp, int p, const NmsWs& ws, int wfirst, int wstride) {
  const int lane = threadIdx.x & 31;
  const int ns = (sp.nb + 31) >> 5;
  for (int sb = wfirst; sb < ns; sb += wstride) {
    const int b = sb * 32 + lane;
    const bool valid = b < sp.nb;
    const double* bb = ws.bbox + 4 * (sp.blk0 + (valid ? b : 0));
    const double x0 = warp_min_d(valid ? __ldcg(bb + 0) : DBL_MAX), y0 = warp_min_d(valid ? __ldcg(bb + 1) : DBL_MAX);
    const double x1 = warp_max_d(valid ? __ldcg(bb + 2) : -DBL_MAX), y1 = warp_max_d(valid ? __ldcg(bb + 3) : -DBL_MAX);
    if (lane == 0) {
      double* o = ws.sbbox + 4 * ((sp.blk0 >> 5) + p + sb);
      o[0] = x0; o[1] = y0; o[2] = x1; o[3] = y1;
    }
  }
}

__global__ void __launch_bounds__(1024) nms_bin_kernel(const double* __restrict__ boxes,
                                                       const double* __restrict__ scores,
                                                       const double* __restrict__ classes,
                                                       const int32_t* __restrict__ sel_idx,
                                                       const int64_t* __restrict__ page_off,
                                                       const int32_t* __restrict__ n_sel, int n_pages, NmsWs ws) {
  __shared__ int hist[32 * NMS_GY];
  __shared__ double red[5][32];
  __shared__ int scan_smem[34];
  const int p = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const PageSpan sp = page_span(page_off, n_sel, p);
  if (blockIdx.x == 0 && tid < 8) ws.stats[tid] = 0;  // status / counters of this call (the first kernel of the call)
  const int64_t blk_next = (p + 1 < n_pages) ? (page_off[p + 1] >> 5) + p + 1 : ws.nb_cap;
  for (int64_t b = sp.blk0 + sp.nb + tid; b < blk_next; b += blockDim.x) ws.blk_page[b] = -1;

  // extent of the box centres and mean box width
  double mnx = DBL_MAX, mny = DBL_MAX, mxx = -DBL_MAX, mxy = -DBL_MAX, sw = 0.0;
  for (int k = tid; k < sp.m; k += blockDim.x) {
    const int64_t gi = sel_idx ? (int64_t)sel_idx[sp.base + k] : sp.base + k;
    const double2 a = *reinterpret_cast<const double2*>(boxes + 4 * gi);
    const double2 b = *reinterpret_cast<const double2*>(boxes + 4 * gi + 2);
    const double cx = (a.x + b.x) * 0.5, cy = (a.y + b.y) * 0.5;
    mnx = fmin(mnx, cx); mxx = fmax(mxx, cx); mny = fmin(mny, cy); mxy = fmax(mxy, cy);
    sw += fabs(b.x - a.x);
  }
  mnx = warp_min_d(mnx); mny = warp_min_d(mny); mxx = warp_max_d(mxx); mxy = warp_max_d(mxy); sw = warp_sum_d(sw);
  if (lane == 0) { red[0][warp] = mnx; red[1][warp] = mny; red[2][warp] = mxx; red[3][warp] = mxy; red[4][warp] = sw; }
  __syncthreads();
  if (warp == 0) {
    const double a = warp_min_d(red[0][lane]), b = warp_min_d(red[1][lane]);
    const double c = warp_max_d(red[2][lane]), d = warp_max_d(red[3][lane]), e = warp_sum_d(red[4][lane]);
    if (lane == 0) { red[0][0] = a; red[1][0] = b; red[2][0] = c; red[3][0] = d; red[4][0] = e; }
  }
  __syncthreads();
  mnx = red[0][0]; mny = red[1][0]; mxx = red[2][0]; mxy = red[3][0];
  // one x strip per mean box width: the boxes of one text column share a strip (pure heuristic:
  // any ordering gives the same result, a good one gives fewer candidate block pairs)
  int gxn = 1;
  if (sp.m > 0 && mxx > mnx) {
    const double meanw = red[4][0] / (double)sp.m;
    const double s = meanw > 0.0 ? (mxx - mnx) / meanw : 1.0;
    gxn = s >= (double)NMS_GX_MAX ? NMS_GX_MAX : (s >= 1.0 ? (int)s + 1 : 1);
    if (gxn > NMS_GX_MAX) gxn = NMS_GX_MAX;
  }
  const double sx = (mxx > mnx) ? (double)gxn / (mxx - mnx) : 0.0;
  const double sy = (mxy > mny) ? (double)NMS_GY / (mxy - mny) : 0.0;

  for (int k = tid; k < sp.m; k += blockDim.x) {
    const int64_t gi = sel_idx ? (int64_t)sel_idx[sp.base + k] : sp.base + k;
    const double2 a = *reinterpret_cast<const double2*>(boxes + 4 * gi);
    const double2 b = *reinterpret_cast<const double2*>(boxes + 4 * gi + 2);
    const double fx = ((a.x + b.x) * 0.5 - mnx) * sx, fy = ((a.y + b.y) * 0.5 - mny) * sy;
    ws.cellid[sp.base + k] = nms_cell_id(fx, fy, gxn);
  }
  __syncthreads();
  // stable LSD radix sort by (x strip, y cell[, y sub-cell]): the finest digit first
  const int32_t* cellid = ws.cellid + sp.base;
  int32_t* tmp = ws.kpos + sp.base;
  int32_t* sorted = ws.sorted_pos + sp.base;
  if (nms_fine_cells(sp.m, gxn, ws.fine_min)) {
    int32_t* tmp0 = reinterpret_cast<int32_t*>(ws.kscore + sp.base);  // free until the resolve kernel writes the kept scores
    block_stable_pass<NMS_GYF>(sp.m, hist, scan_smem, [](int i) { return i; },
                               [cellid](int e) { return cellid[e] & (NMS_GYF - 1); }, tmp0);
    block_stable_pass<NMS_GY>(sp.m, hist, scan_smem, [tmp0](int i) { return tmp0[i]; },
                              [cellid](int e) { return (cellid[e] >> 7) & (NMS_GY - 1); }, tmp);
  } else {
    block_stable_pass<NMS_GY>(sp.m, hist, scan_smem, [](int i) { return i; },
                              [cellid](int e) { return (cellid[e] >> 7) & (NMS_GY - 1); }, tmp);
  }
  block_stable_pass<NMS_GX_MAX>(sp.m, hist, scan_smem, [tmp](int i) { return tmp[i]; },
                                [cellid](int e) { return cellid[e] >> 14; }, sorted);

  // blocked copy + block bounding boxes
  nms_blocked_copy(boxes, scores, classes, sel_idx, sp, sorted, ws, p, warp, blockDim.x >> 5);
  __threadfence_block();
  __syncthreads();
  nms_super_bounds(sp, p, ws, warp, blockDim.x >> 5);
}

// Blocked AoS copy of a page in spatial order + block / sub-block bounding boxes; block b is handled by the
// warp whose index is congruent to b modulo `wstride` (the warps of one CTA, or of all CTAs of a cluster).
__device__ __forceinline__ void nms_blocked_copy(const double* __restrict__ boxes, const double* __restrict__ scores,
                                                 const double* __restrict__ classes, const int32_t* __restrict__ sel_idx,
                                                 const PageSpan& sp, const int32_t* sorted, const NmsWs& ws, int p,
                                                 int wfirst, int wstride) {
  const int lane = threadIdx.x & 31;
  for (int b = wfirst; b < sp.nb; b += wstride) {
    const int pos = b * 32 + lane;
    const bool valid = pos < sp.m;
    SBox sb;
    sb.x0 = sb.y0 = sb.x1 = sb.y1 = sb.area = sb.score = sb.cls = 0.0;
    sb.k = -1;
    if (valid) {
      const int k = sorted[pos];
      PG_DEV_ASSERT(k >= 0 && k < sp.m);
      const int64_t gi = sel_idx ? (int64_t)sel_idx[sp.base + k] : sp.base + k;
      const double2 a = *reinterpret_cast<const double2*>(boxes + 4 * gi);
      const double2 c = *reinterpret_cast<const double2*>(boxes + 4 * gi + 2);
      sb.x0 = a.x; sb.y0 = a.y; sb.x1 = c.x; sb.y1 = c.y;
      sb.area = (ws.mode & PG_NMS_FP32) ? (double)pg_box_area_f32((float)a.x, (float)a.y, (float)c.x, (float)c.y)
                                        : pg_box_area(a.x, a.y, c.x, c.y);
      sb.score = scores[gi]; sb.cls = classes ? classes[gi] : 0.0; sb.k = k;
    }
    ws.sbox[(sp.blk0 + b) * 32 + lane] = sb;
    const double bx0 = warp_min_d(valid ? fmin(sb.x0, sb.x1) : DBL_MAX), by0 = warp_min_d(valid ? fmin(sb.y0, sb.y1) : DBL_MAX);
    const double bx1 = warp_max_d(valid ? fmax(sb.x0, sb.x1) : -DBL_MAX), by1 = warp_max_d(valid ? fmax(sb.y0, sb.y1) : -DBL_MAX);
    if (lane == 0) {
      double* bb = ws.bbox + 4 * (sp.blk0 + b);
      bb[0] = bx0; bb[1] = by0; bb[2] = bx1; bb[3] = by1;
      ws.bboxf[sp.blk0 + b] = make_float4(__double2float_rd(bx0), __double2float_rd(by0), __double2float_ru(bx1),
                                          __double2float_ru(by1));
      ws.blk_page[sp.blk0 + b] = p;
    }
    // bounds of each run of 8 boxes, rounded outwards: lets the mask kernel skip 8 pair tests at a time
    float gx0 = valid ? __double2float_rd(fmin(sb.x0, sb.x1)) : FLT_MAX, gy0 = valid ? __double2float_rd(fmin(sb.y0, sb.y1)) : FLT_MAX;
    float gx1 = valid ? __double2float_ru(fmax(sb.x0, sb.x1)) : -FLT_MAX, gy1 = valid ? __double2float_ru(fmax(sb.y0, sb.y1)) : -FLT_MAX;
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
      gx0 = fminf(gx0, __shfl_xor_sync(0xffffffffu, gx0, o)); gy0 = fminf(gy0, __shfl_xor_sync(0xffffffffu, gy0, o));
      gx1 = fmaxf(gx1, __shfl_xor_sync(0xffffffffu, gx1, o)); gy1 = fmaxf(gy1, __shfl_xor_sync(0xffffffffu, gy1, o));
    }
    if ((lane & 7) == 0) ws.subbox[(sp.blk0 + b) * 4 + (lane >> 3)] = make_float4(gx0, gy0, gx1, gy1);
  }
}

// ---- A (large pages): bin on one thread-block cluster per page --------------------------------
// One stable counting-sort pass over the cluster: global warp g = rank * 32 + warp owns a contiguous chunk
// of the input, so (digit, g, position) order == (digit, input order).  Each CTA publishes its per-digit
// totals to every CTA of the cluster through distributed shared memory; the bases follow locally.
template <int NDIG, typename ElemOf, typename DigitOf>
__device__ __forceinline__ void cluster_stable_pass(cg::cluster_group& cluster, int csize, int rank, int m, int* hist,
                                                    int* all_tot, int* scan_smem, ElemOf elem, DigitOf digit,
                                                    int32_t* dst) {
  static_assert(NDIG <= 1024, "one thread per digit");
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < 32 * NDIG; i += 1024) hist[i] = 0;
  __syncthreads();
  const int tw = csize * 32;
  const int chunk = (((m + tw - 1) / tw) + 31) & ~31;
  const int lo = min(m, (rank * 32 + warp) * chunk), hi = min(m, lo + chunk);
  for (int i = lo + lane; i < hi; i += 32) atomicAdd(&hist[warp * NDIG + digit(elem(i))], 1);
  __syncthreads();
  if (tid < NDIG) {
    int t = 0;
    for (int w = 0; w < 32; ++w) t += hist[w * NDIG + tid];
    for (int q = 0; q < csize; ++q) *cluster.map_shared_rank(&all_tot[rank * NDIG + tid], q) = t;
  }
  cluster.sync();
  {
    int v = 0, before = 0;
    if (tid < NDIG) {
      for (int q = 0; q < csize; ++q) {
        const int t = all_tot[q * NDIG + tid];
        v += t;
        if (q < rank) before += t;
      }
    }
    int total;
    const int ex = pg_block_exscan(v, scan_smem, &total);  // digit bases of the whole page
    if (tid < NDIG) {
      int running = ex + before;
      for (int w = 0; w < 32; ++w) {
        const int t = hist[w * NDIG + tid];
        hist[w * NDIG + tid] = running;
        running += t;
      }
    }
  }
  __syncthreads();
  for (int c = lo; c < hi; c += 32) {
    const int i = c + lane;
    const bool act = i < hi;
    const int e = act ? elem(i) : 0;
    const int d = act ? digit(e) : -1 - lane;
    const unsigned grp = __match_any_sync(0xffffffffu, d);
    const int leader = __ffs(grp) - 1;
    const int rk = __popc(grp & ((1u << lane) - 1u));
    int basepos = 0;
    if (act && lane == leader) {
      basepos = hist[warp * NDIG + d];
      hist[warp * NDIG + d] = basepos + __popc(grp);
    }
    basepos = __shfl_sync(0xffffffffu, basepos, leader);
    PG_DEV_ASSERT(!act || (basepos + rk >= 0 && basepos + rk < m && e >= 0 && e < m));
    if (act) dst[basepos + rk] = e;
    __syncwarp();
  }
  __threadfence();
  cluster.sync();  // dst complete and visible to the whole cluster; all_tot free for the next pass
}

__global__ void __launch_bounds__(1024) nms_bin_cluster_kernel(const double* __restrict__ boxes,
                                                               const double* __restrict__ scores,
                                                               const double* __restrict__ classes,
                                                               const int32_t* __restrict__ sel_idx,
                                                               const int64_t* __restrict__ page_off,
                                                               const int32_t* __restrict__ n_sel, int n_pages, NmsWs ws) {
  __shared__ int hist[32 * NMS_GY];
  __shared__ int all_tot[NMS_CLUSTER_MAX * NMS_GY];
  __shared__ double red[5][32];
  __shared__ double ext_x[NMS_CLUSTER_MAX][5];
  __shared__ int scan_smem[34];
  static_assert(NMS_GX_MAX <= NMS_GY, "hist / all_tot are sized for the wider digit");
  cg::cluster_group cluster = cg::this_cluster();
  const int csize = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
  const int p = blockIdx.x / csize, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ctid = rank * (int)blockDim.x + tid, cthreads = csize * (int)blockDim.x;
  const PageSpan sp = page_span(page_off, n_sel, p);
  if (blockIdx.x == 0 && tid < 8) ws.stats[tid] = 0;  // status / counters of this call (the first kernel of the call)
  const int64_t blk_next = (p + 1 < n_pages) ? (page_off[p + 1] >> 5) + p + 1 : ws.nb_cap;
  for (int64_t b = sp.blk0 + sp.nb + ctid; b < blk_next; b += cthreads) ws.blk_page[b] = -1;

  // extent of the box centres and mean box width
  double mnx = DBL_MAX, mny = DBL_MAX, mxx = -DBL_MAX, mxy = -DBL_MAX, sw = 0.0;
  for (int k = ctid; k < sp.m; k += cthreads) {
    const int64_t gi = sel_idx ? (int64_t)sel_idx[sp.base + k] : sp.base + k;
    const double2 a = *reinterpret_cast<const double2*>(boxes + 4 * gi);
    const double2 b = *reinterpret_cast<const double2*>(boxes + 4 * gi + 2);
    const double cx = (a.x + b.x) * 0.5, cy = (a.y + b.y) * 0.5;
    mnx = fmin(mnx, cx); mxx = fmax(mxx, cx); mny = fmin(mny, cy); mxy = fmax(mxy, cy);
    sw += fabs(b.x - a.x);
  }
  mnx = warp_min_d(mnx); mny = warp_min_d(mny); mxx = warp_max_d(mxx); mxy = warp_max_d(mxy); sw = warp_sum_d(sw);
  if (lane == 0) { red[0][warp] = mnx; red[1][warp] = mny; red[2][warp] = mxx; red[3][warp] = mxy; red[4][warp] = sw; }
  __syncthreads();
  if (warp == 0) {
    const double a = warp_min_d(red[0][lane]), b = warp_min_d(red[1][lane]);
    const double c = warp_max_d(red[2][lane]), d = warp_max_d(red[3][lane]), e = warp_sum_d(red[4][lane]);
    if (lane == 0) { red[0][0] = a; red[1][0] = b; red[2][0] = c; red[3][0] = d; red[4][0] = e; }
  }
  __syncthreads();
  if (tid < 5 * csize) *cluster.map_shared_rank(&ext_x[rank][tid % 5], tid / 5) = red[tid % 5][0];
  cluster.sync();
  mnx = ext_x[0][0]; mny = ext_x[0][1]; mxx = ext_x[0][2]; mxy = ext_x[0][3];
  double sw_all = ext_x[0][4];
  for (int q = 1; q < csize; ++q) {  // the same order in every CTA: all of them derive the same grid
    mnx = fmin(mnx, ext_x[q][0]); mny = fmin(mny, ext_x[q][1]);
    mxx = fmax(mxx, ext_x[q][2]); mxy = fmax(mxy, ext_x[q][3]);
    sw_all += ext_x[q][4];
  }
  int gxn = 1;
  if (sp.m > 0 && mxx > mnx) {
    const double meanw = sw_all / (double)sp.m;
    const double s = meanw > 0.0 ? (mxx - mnx) / meanw : 1.0;
    gxn = s >= (double)NMS_GX_MAX ? NMS_GX_MAX : (s >= 1.0 ? (int)s + 1 : 1);
    if (gxn > NMS_GX_MAX) gxn = NMS_GX_MAX;
  }
  const double sx = (mxx > mnx) ? (double)gxn / (mxx - mnx) : 0.0;
  const double sy = (mxy > mny) ? (double)NMS_GY / (mxy - mny) : 0.0;

  for (int k = ctid; k < sp.m; k += cthreads) {
    const int64_t gi = sel_idx ? (int64_t)sel_idx[sp.base + k] : sp.base + k;
    const double2 a = *reinterpret_cast<const double2*>(boxes + 4 * gi);
    const double2 b = *reinterpret_cast<const double2*>(boxes + 4 * gi + 2);
    const double fx = ((a.x + b.x) * 0.5 - mnx) * sx, fy = ((a.y + b.y) * 0.5 - mny) * sy;
    ws.cellid[sp.base + k] = nms_cell_id(fx, fy, gxn);
  }
  __threadfence();
  cluster.sync();  // the passes read cell ids written by other CTAs
  const int32_t* cellid = ws.cellid + sp.base;
  int32_t* tmp = ws.kpos + sp.base;
  int32_t* sorted = ws.sorted_pos + sp.base;
  if (nms_fine_cells(sp.m, gxn, ws.fine_min)) {  // the same decision in every CTA of the cluster
    int32_t* tmp0 = reinterpret_cast<int32_t*>(ws.kscore + sp.base);
    cluster_stable_pass<NMS_GYF>(cluster, csize, rank, sp.m, hist, all_tot, scan_smem, [](int i) { return i; },
                                 [cellid](int e) { return cellid[e] & (NMS_GYF - 1); }, tmp0);
    cluster_stable_pass<NMS_GY>(cluster, csize, rank, sp.m, hist, all_tot, scan_smem, [tmp0](int i) { return tmp0[i]; },
                                [cellid](int e) { return (cellid[e] >> 7) & (NMS_GY - 1); }, tmp);
  } else {
    cluster_stable_pass<NMS_GY>(cluster, csize, rank, sp.m, hist, all_tot, scan_smem, [](int i) { return i; },
                                [cellid](int e) { return (cellid[e] >> 7) & (NMS_GY - 1); }, tmp);
  }
  cluster_stable_pass<NMS_GX_MAX>(cluster, csize, rank, sp.m, hist, all_tot, scan_smem, [tmp](int i) { return tmp[i]; },
                                  [cellid](int e) { return cellid[e] >> 14; }, sorted);
  nms_blocked_copy(boxes, scores, classes, sel_idx, sp, sorted, ws, p, rank * 32 + warp, csize * 32);
  __threadfence();
  cluster.sync();  // the block bounds of the page were written by all CTAs of the cluster
  nms_super_bounds(sp, p, ws, rank * 32 + warp, csize * 32);
}

__device__ __forceinline__ bool bbox_hit(const double* a, const double* b) {
  return !(b[2] < a[0] || a[2] < b[0] || b[3] < a[1] || a[3] < b[1]);
}

// One warp per block I: count the blocks J of the page whose bounding boxes intersect bbox(I),
// reserve that many entries with one atomicAdd, then list them (entries of one I stay contiguous and
// ascending in J; their global order depends on the atomics but nothing downstream depends on it).
// Two levels: the super-blocks of the page first, then the 32 blocks of each super-block that is hit.
template <typename OnHit>
__device__ __forceinline__ void cand_walk(const NmsWs& ws, const PageSpan& sp, int p, const double* bbI, int all_pairs,
                                          int lane, OnHit on_hit) {
  const int ns = (sp.nb + 31) >> 5;
  const double* sb = ws.sbbox + 4 * ((sp.blk0 >> 5) + p);
  for (int s0 = 0; s0 < ns; s0 += 32) {
    const bool sh = (s0 + lane < ns) && (all_pairs || bbox_hit(bbI, sb + 4 * (s0 + lane)));
    unsigned sm = __ballot_sync(0xffffffffu, sh);
    while (sm) {
      const int j = (s0 + __ffs(sm) - 1) * 32 + lane;
      sm &= sm - 1;
      const bool hit = (j < sp.nb) && (all_pairs || bbox_hit(bbI, ws.bbox + 4 * (sp.blk0 + j)));
      on_hit(j, hit, __ballot_sync(0xffffffffu, hit));
    }
  }
}

__global__ void __launch_bounds__(256) nms_cand_kernel(const int64_t* __restrict__ page_off,
                                                       const int32_t* __restrict__ n_sel, NmsWs ws, int all_pairs) {
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t I = (int64_t)blockIdx.x * (blockDim.x >> 5) + wib;
  if (I >= ws.nb_cap) return;
  const int p = ws.blk_page[I];
  if (p < 0) {
    if (lane == 0) { ws.cand_cnt[I] = 0; ws.cand_off[I] = 0; }
    return;
  }
  const PageSpan sp = page_span(page_off, n_sel, p);
  double bbI[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) bbI[q] = ws.bbox[4 * I + q];
  // The counting walk also remembers the first CAND_KEEP hits of the block in shared memory: the listing below then
  // copies them instead of walking the page's bounding boxes a second time (cfg4: 11 candidates per block on average).
  constexpr int CAND_KEEP = 96;
  __shared__ int32_t kept_j[8][CAND_KEEP];
  int cnt = 0;
  cand_walk(ws, sp, p, bbI, all_pairs, lane, [&](int j, bool hit, unsigned hits) {
    if (hit) {
      const int slot = cnt + __popc(hits & ((1u << lane) - 1u));
      if (slot < CAND_KEEP) kept_j[wib][slot] = j;
    }
    cnt += __popc(hits);
  });
  __syncwarp();
  long long off = 0;
  if (lane == 0) off = (long long)atomicAdd((unsigned long long*)&ws.stats[ST_ENT_TOTAL], (unsigned long long)cnt);
  off = __shfl_sync(0xffffffffu, off, 0);
  const bool fits = off + cnt <= ws.ent_cap;
  if (lane == 0) {
    ws.cand_off[I] = off;
    ws.cand_cnt[I] = fits ? cnt : 0;
    if (!fits) ws.stats[ST_STATUS] = PG_ERR_WORKSPACE;  // every writer stores the same value
  }
  if (!fits) return;
  if (cnt <= CAND_KEEP) {
    for (int k = lane; k < cnt; k += 32) {
      PG_DEV_ASSERT(off + k < ws.ent_cap && kept_j[wib][k] >= 0 && kept_j[wib][k] < sp.nb);
      ws.ent_j[off + k] = (int32_t)(sp.blk0 + kept_j[wib][k]);
      ws.ent_i[off + k] = (int32_t)I;
    }
    return;
  }
  long long e = off;
  cand_walk(ws, sp, p, bbI, all_pairs, lane, [&](int j, bool hit, unsigned hits) {
    if (hit) {
      const long long slot = e + __popc(hits & ((1u << lane) - 1u));
      PG_DEV_ASSERT(slot >= 0 && slot < ws.ent_cap && j >= 0 && j < sp.nb);
      ws.ent_j[slot] = (int32_t)(sp.blk0 + j);
      ws.ent_i[slot] = (int32_t)I;
    }
    e += __popc(hits);
  });
}

// ---- C: masks ------------------------------------------------------------------------------
// Persistent grid over candidate entries (each entry = 32x32 box pairs, so the work is balanced
// whatever the spread of candidates per block).  Lane i holds box i of block I; the boxes of J sit in
// shared memory twice: as 32-byte prefilter records (outward-rounded fp32 bounds, score, position,
// class hash — broadcast reads, 2x LDS.128 per box) and in full.  Phase 1 marks, per lane, the boxes
// of J that outrank box i (:112), may share its class (:130) and are not provably disjoint from it
// (:65); runs of 8 boxes of J whose bounds miss block I are skipped warp-wide.  Phase 2 evaluates the
// exact predicate (exact class compare, pg_iou_gt) on the marked pairs only.  The kernel is bound by
// shared-memory bandwidth, hence the compact records and the plane layout of the full boxes (the AoS
// layout cost 4-way bank conflicts on every store and on the lane-indexed reads of phase 2).
constexpr int MASK_UNIT = 4;  // consecutive entries per work unit (mostly the same I)
struct __align__(16) SLite {
  float x0, y0, x1, y1;  // box rounded outwards: disjoint here => disjoint in fp64
  double score;
  int k;                 // pooled position, -1 = padding lane
  uint32_t ch;           // class hash: equal classes => equal hashes
};
__device__ __forceinline__ SLite slite_of(const SBox& b) {
  SLite s;
  s.x0 = __double2float_rd(fmin(b.x0, b.x1)); s.y0 = __double2float_rd(fmin(b.y0, b.y1));
  s.x1 = __double2float_ru(fmax(b.x0, b.x1)); s.y1 = __double2float_ru(fmax(b.y0, b.y1));
  s.score = b.score;
  s.k = (int)b.k;
  const double c = (b.cls == 0.0) ? 0.0 : b.cls;  // -0.0 == +0.0 must hash alike
  const unsigned long long u = (unsigned long long)__double_as_longlong(c);
  s.ch = (uint32_t)(u >> 32) ^ (uint32_t)u;
  return s;
}

// One prefilter test of the mask kernel as a single predicate chain (nine instructions, no branch, no
// materialised booleans): box j outranks box i (3_combine_grids.py:112: higher score, earlier position on ties),
// carries the same class hash (:130) and is not provably disjoint from it (:65; `!(a < b)` keeps the
// unordered case, like the C expression).  Returns bits | bit when all of that holds.
__device__ __forceinline__ uint32_t prefilter_or(uint32_t bits, uint32_t bit, double sj, double si, int kj, int ki,
                                                 uint32_t chj, uint32_t chi, float4 bj, float ix0, float iy0, float ix1,
                                                 float iy1) {
  asm("{\n\t"
      ".reg .pred p;\n\t"
      "setp.eq.f64 p, %2, %3;\n\t"
      "setp.lt.and.s32 p, %4, %5, p;\n\t"
      "setp.gt.or.f64 p, %2, %3, p;\n\t"
      "setp.eq.and.u32 p, %6, %7, p;\n\t"
      "setp.geu.and.f32 p, %8, %9, p;\n\t"    // !(bj.z < ix0)
      "setp.geu.and.f32 p, %10, %11, p;\n\t"  // !(ix1 < bj.x)
      "setp.geu.and.f32 p, %12, %13, p;\n\t"  // !(bj.w < iy0)
      "setp.geu.and.f32 p, %14, %15, p;\n\t"  // !(iy1 < bj.y)
      "@p or.b32 %0, %0, %1;\n\t"
      "}"
      : "+r"(bits)
      : "r"(bit), "d"(sj), "d"(si), "r"(kj), "r"(ki), "r"(chj), "r"(chi), "f"(bj.z), "f"(ix0), "f"(ix1), "f"(bj.x),
        "f"(bj.w), "f"(iy0), "f"(iy1), "f"(bj.y));
  return bits;
}

template <int MIN_CTAS>
__global__ void __launch_bounds__(256, MIN_CTAS) nms_mask_kernel(NmsWs ws, double thr) {
  // boxes of J: full-precision fields as planes (8-byte words, box index innermost: stores and the
  // lane-indexed reads of phase 2 are free of bank conflicts), prefilter records as two 16-byte halves
  __shared__ double jbf[8][6][32];  // x0, y0, x1, y1, area, cls
  __shared__ float4 jla[8][32];     // outward-rounded fp32 bounds (all of the plane when nothing may be skipped)
  __shared__ int4 jlb[8][32];       // score (lo, hi), position, class hash (0 when classes are ignored)
  if (ws.stats[ST_STATUS] != PG_OK) return;
  const long long total = ws.stats[ST_ENT_TOTAL];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const long long gw = (long long)blockIdx.x * 8 + wib, nw = (long long)gridDim.x * 8;
  const bool all_pairs = !(thr >= 0.0);  // thr < 0: IoU 0 already suppresses, nothing can be skipped
  const bool agnostic = (ws.mode & PG_NMS_CLASS_AGNOSTIC) != 0, f32 = (ws.mode & PG_NMS_FP32) != 0;
  // the two launch-wide switches are folded into the records, so that the 1024 prefilter tests of an entry
  // are plain compare chains: class hashes masked to 0 when classes are ignored, bounds opened to the whole
  // plane when every same-class pair has to be tested
  const uint32_t ch_keep = agnostic ? 0u : 0xffffffffu;
  int cached = -1;
  SBox bi;
  bi.k = -1;
  SLite li = slite_of(bi);
  uint32_t li_ch = 0u;
  float4 bbI = make_float4(0.f, 0.f, 0.f, 0.f);
  unsigned long long runs_tested = 0;  // runs of 8 boxes of J whose prefilter chains were actually evaluated (x 256 pairs)
  for (long long u = gw; u * MASK_UNIT < total; u += nw) {
    for (int q = 0; q < MASK_UNIT; ++q) {
      const long long e = u * MASK_UNIT + q;
      if (e >= total) break;
      const int I = ws.ent_i[e], J = ws.ent_j[e];
      PG_DEV_ASSERT(e < ws.ent_cap && I >= 0 && I < ws.nb_cap && J >= 0 && J < ws.nb_cap);
      if (I != cached) {
        bi = ws.sbox[(int64_t)I * 32 + lane];
        li = slite_of(bi);
        li_ch = li.ch & ch_keep;
        bbI = ws.bboxf[I];
        cached = I;
      }
      __syncwarp();
      {
        const SBox bj = ws.sbox[(int64_t)J * 32 + lane];
        const SLite lj = slite_of(bj);
        jbf[wib][0][lane] = bj.x0; jbf[wib][1][lane] = bj.y0; jbf[wib][2][lane] = bj.x1; jbf[wib][3][lane] = bj.y1;
        jbf[wib][4][lane] = bj.area; jbf[wib][5][lane] = bj.cls;
        // padding lanes of J are disjoint from everything, real boxes from nothing when all pairs must be tested
        jla[wib][lane] = lj.k < 0 ? make_float4(INFINITY, INFINITY, -INFINITY, -INFINITY)
                                  : (all_pairs ? make_float4(-INFINITY, -INFINITY, INFINITY, INFINITY)
                                               : make_float4(lj.x0, lj.y0, lj.x1, lj.y1));
        jlb[wib][lane] = make_int4(__double2loint(lj.score), __double2hiint(lj.score), lj.k, (int)(lj.ch & ch_keep));
      }
      // which runs of 8 boxes of J can touch block I at all (warp-uniform)
      bool ghit = false;
      if (lane < 4) {
        const float4 sb = ws.subbox[(int64_t)J * 4 + lane];
        ghit = all_pairs || !(sb.z < bbI.x || bbI.z < sb.x || sb.w < bbI.y || bbI.w < sb.y);
      }
      const unsigned groups = __ballot_sync(0xffffffffu, ghit) & 0xFu;
      runs_tested += (unsigned)__popc(groups);
      __syncwarp();
      // phase 1
      uint32_t cand = 0;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        if (groups & (1u << g)) {
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            const int jj = g * 8 + t;
            const float4 la = jla[wib][jj];  // broadcast LDS.128 x2
            const int4 lb = jlb[wib][jj];
            cand = prefilter_or(cand, 1u << jj, __hiloint2double(lb.y, lb.x), li.score, lb.z, li.k, (uint32_t)lb.w, li_ch,
                                la, li.x0, li.y0, li.x1, li.y1);
          }
        }
      }
      if (li.k < 0) cand = 0u;  // padding lane of I
      // phase 2 (exact, only on the marked pairs; lanes walk their own bit lists)
      uint32_t mask = 0;
      while (__any_sync(0xffffffffu, cand != 0)) {
        if (cand) {
          const int jj = __ffs(cand) - 1;
          cand &= cand - 1;
          if (agnostic || jbf[wib][5][jj] == bi.cls) {
            const double jx0 = jbf[wib][0][jj], jy0 = jbf[wib][1][jj], jx1 = jbf[wib][2][jj], jy1 = jbf[wib][3][jj];
            const double jarea = jbf[wib][4][jj];
            const bool hit = f32 ? pg_iou_gt_f32((float)jx0, (float)jy0, (float)jx1, (float)jy1, (float)jarea,
                                                 (float)bi.x0, (float)bi.y0, (float)bi.x1, (float)bi.y1, (float)bi.area, thr)
                                 : pg_iou_gt(jx0, jy0, jx1, jy1, jarea, bi.x0, bi.y0, bi.x1, bi.y1, bi.area, thr);
            if (hit) mask |= 1u << jj;
          }
        }
      }
      const bool any = __any_sync(0xffffffffu, mask != 0);
      ws.ent_mask[e * 32 + lane] = mask;
      if (lane == 0 && !any) ws.ent_j[e] = -1;
    }
  }
  // box pairs that went through the prefilter: counted, not nominal (runs skipped by their bounds are not in it)
  if (lane == 0 && runs_tested) atomicAdd((unsigned long long*)&ws.stats[ST_TESTS], runs_tested * 256ull);
}

// ---- E: resolve ----------------------------------------------------------------------------
// The candidate entries of block I against the state words of round `rc`: box `lane` of I is suppressed
// if a kept box of some J suppresses it, and has to wait while an undecided one might.  The entry
// headers (block id and the two state words of J) are fetched 32 at a time, one per lane; only entries
// whose J still has kept or undecided boxes touch their 128-byte mask row, four rows in flight.
__device__ __forceinline__ void resolve_scan_entries(const NmsWs& ws, const volatile uint32_t* kept,
                                                     const volatile uint32_t* undec, int64_t rc, int64_t I, bool mine,
                                                     int lane, bool& sup, bool& wait) {
  const int64_t e0 = ws.cand_off[I], e1 = e0 + ws.cand_cnt[I];
  PG_DEV_ASSERT(e0 >= 0 && e1 <= ws.ent_cap);
  for (int64_t eb = e0; eb < e1; eb += 32) {
    const int64_t me = eb + lane;
    const int jbk = me < e1 ? ws.ent_j[me] : -1;  // global block id, -1 = no suppressor in that block
    PG_DEV_ASSERT(jbk < ws.nb_cap);
    uint32_t kj = 0u, uj = 0u;
    if (jbk >= 0) { kj = kept[rc + jbk]; uj = undec[rc + jbk]; }
    unsigned live = __ballot_sync(0xffffffffu, (kj | uj) != 0u);
    while (live) {
      int q[4];
      uint32_t m[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        q[t] = live ? __ffs(live) - 1 : -1;
        live &= live - 1;  // 0 stays 0
      }
#pragma unroll
      for (int t = 0; t < 4; ++t) m[t] = (mine && q[t] >= 0) ? ws.ent_mask[(eb + q[t]) * 32 + lane] : 0u;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const uint32_t kq = __shfl_sync(0xffffffffu, kj, q[t] & 31), uq = __shfl_sync(0xffffffffu, uj, q[t] & 31);
        if (m[t] & kq) sup = true;
        else if (m[t] & uq) wait = true;
      }
      if (__all_sync(0xffffffffu, !mine || sup)) return;  // every undecided box of I is suppressed already
    }
  }
}

// One Jacobi round over the blocks first, first + stride, ... a warp owns.  A block that was final in the previous
// round already is copied once into the other buffer and then leaves the warp's `alive` mask (bit i = i-th owned
// block): later rounds visit only blocks that can still change.  Warps that own more than 32 blocks keep visiting
// all of them.  Returns whether any visited block is still undecided.
__device__ __forceinline__ int resolve_round(const NmsWs& ws, volatile uint32_t* kept, volatile uint32_t* undec, int64_t rc,
                                             int64_t rn, const PageSpan& sp, int first, int stride, int owned, uint32_t& alive,
                                             int lane) {
  int pending = 0;
  auto visit = [&](int b) -> bool {
    const int64_t I = sp.blk0 + b;
    const uint32_t u = undec[rc + I], k = kept[rc + I];
    if (u == 0u) {
      if (lane == 0) { undec[rn + I] = 0u; kept[rn + I] = k; }
      return false;
    }
    const bool mine = (u >> lane) & 1u;
    bool sup = false, wait = false;
    resolve_scan_entries(ws, kept, undec, rc, I, mine, lane, sup, wait);
    const unsigned bk = __ballot_sync(0xffffffffu, mine && !sup && !wait);
    const unsigned bs = __ballot_sync(0xffffffffu, mine && sup);
    const uint32_t un = u & ~(bk | bs);
    if (lane == 0) { kept[rn + I] = k | bk; undec[rn + I] = un; }
    pending |= (un != 0u);
    return true;
  };
  if (owned <= 32) {
    for (uint32_t am = alive; am; am &= am - 1u) {
      const int i = __ffs(am) - 1;
      if (!visit(first + i * stride)) alive &= ~(1u << i);
    }
  } else {
    for (int b = first; b < sp.nb; b += stride) visit(b);
  }
  return pending;
}

__global__ void __launch_bounds__(1024) nms_resolve_kernel(const int64_t* __restrict__ page_off,
                                                           const int32_t* __restrict__ n_sel, NmsWs ws,
                                                           int32_t* __restrict__ n_kept) {
  __shared__ int scan_smem[34];
  const int p = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const PageSpan sp = page_span(page_off, n_sel, p);
  if (ws.stats[0] != PG_OK) {
    if (tid == 0) n_kept[p] = -1;
    return;
  }
  // state words are written and re-read by this CTA only; volatile keeps them out of the nc path
  volatile uint32_t* kept = ws.st_kept;
  volatile uint32_t* undec = ws.st_undec;
  const int64_t nbc = ws.nb_cap;
  for (int b = tid; b < sp.nb; b += blockDim.x) {
    const int valid = min(32, sp.m - b * 32);
    undec[sp.blk0 + b] = valid >= 32 ? 0xffffffffu : ((1u << valid) - 1u);
    kept[sp.blk0 + b] = 0u;
  }
  __syncthreads();
  const int owned = warp < sp.nb ? (sp.nb - warp + nwarps - 1) / nwarps : 0;
  uint32_t alive = owned >= 32 ? 0xffffffffu : ((1u << owned) - 1u);
  int cur = 0, rounds = 0;
  while (true) {
    const int64_t rc = (int64_t)cur * nbc, rn = (int64_t)(cur ^ 1) * nbc;
    const int pending = resolve_round(ws, kept, undec, rc, rn, sp, warp, nwarps, owned, alive, lane);
    cur ^= 1;
    ++rounds;
    if (!__syncthreads_or(pending)) break;
  }
  // compact the kept boxes (spatial order) into (score, position) lists
  const int64_t rc = (int64_t)cur * nbc;
  int running = 0;
  for (int b0 = 0; b0 < sp.nb; b0 += blockDim.x) {
    const int b = b0 + tid;
    const uint32_t k = b < sp.nb ? kept[rc + sp.blk0 + b] : 0u;
    int total;
    const int ex = pg_block_exscan(__popc(k), scan_smem, &total);
    uint32_t bits = k;
    int64_t dst = sp.base + running + ex;
    while (bits) {
      const int l = __ffs(bits) - 1;
      bits &= bits - 1;
      const SBox* sb = ws.sbox + (sp.blk0 + b) * 32 + l;
      PG_DEV_ASSERT(dst >= sp.base && dst < sp.base + sp.m);
      ws.kscore[dst] = sb->score;
      ws.kpos[dst] = (int32_t)sb->k;
      ++dst;
    }
    running += total;
  }
  if (tid == 0) {
    n_kept[p] = running;
    atomicMax((unsigned long long*)&ws.stats[2], (unsigned long long)rounds);
  }
}

// ---- E (large pages): resolve on one thread-block cluster per page ---------------------------
// The same Jacobi rounds, the page's blocks dealt round-robin to the warps of all CTAs of the cluster.
// A round reads only the state words of the previous round (double-buffered), so one cluster barrier
// per round orders everything; the state lives in global memory behind volatile (L1-bypassing)
// accesses, the per-round "anything still undecided" flags and the final per-CTA counts travel
// through distributed shared memory.
__global__ void __launch_bounds__(1024) nms_resolve_cluster_kernel(const int64_t* __restrict__ page_off,
                                                                   const int32_t* __restrict__ n_sel, NmsWs ws,
                                                                   int32_t* __restrict__ n_kept) {
  __shared__ int scan_smem[34];
  __shared__ int pend_x[2][NMS_CLUSTER_MAX];  // [round parity][CTA rank]
  __shared__ int tot_x[NMS_CLUSTER_MAX];
  cg::cluster_group cluster = cg::this_cluster();
  const int csize = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
  const int p = blockIdx.x / csize, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const PageSpan sp = page_span(page_off, n_sel, p);
  if (ws.stats[0] != PG_OK) {  // the same word for every CTA of the cluster: all of them leave here
    if (rank == 0 && tid == 0) n_kept[p] = -1;
    return;
  }
  volatile uint32_t* kept = ws.st_kept;
  volatile uint32_t* undec = ws.st_undec;
  const int64_t nbc = ws.nb_cap;
  for (int b = rank * (int)blockDim.x + tid; b < sp.nb; b += csize * (int)blockDim.x) {
    const int valid = min(32, sp.m - b * 32);
    undec[sp.blk0 + b] = valid >= 32 ? 0xffffffffu : ((1u << valid) - 1u);
    kept[sp.blk0 + b] = 0u;
  }
  __threadfence();
  cluster.sync();
  const int first = rank * nwarps + warp;
  const int owned = first < sp.nb ? (sp.nb - first + csize * nwarps - 1) / (csize * nwarps) : 0;
  uint32_t alive = owned >= 32 ? 0xffffffffu : ((1u << owned) - 1u);
  int cur = 0, rounds = 0;
  while (true) {
    const int64_t rc = (int64_t)cur * nbc, rn = (int64_t)(cur ^ 1) * nbc;
    const int pending = resolve_round(ws, kept, undec, rc, rn, sp, first, csize * nwarps, owned, alive, lane);
    const int any_local = __syncthreads_or(pending);
    const int par = rounds & 1;
    if (tid < csize) *cluster.map_shared_rank(&pend_x[par][rank], tid) = any_local;
    __threadfence();
    cluster.sync();
    int any = 0;
    for (int q = 0; q < csize; ++q) any |= pend_x[par][q];
    cur ^= 1;
    ++rounds;
    if (!any) break;
  }
  // compact the kept boxes into (score, position) lists: CTA r owns a contiguous range of blocks
  const int64_t rc = (int64_t)cur * nbc;
  const int per = (sp.nb + csize - 1) / csize;
  const int lo = min(sp.nb, rank * per), hi = min(sp.nb, lo + per);
  int cnt = 0;
  for (int b = lo + tid; b < hi; b += blockDim.x) cnt += __popc(kept[rc + sp.blk0 + b]);
  int mine_total;
  pg_block_exscan(cnt, scan_smem, &mine_total);
  if (tid < csize) *cluster.map_shared_rank(&tot_x[rank], tid) = mine_total;
  cluster.sync();
  int running = 0, all = 0;
  for (int q = 0; q < csize; ++q) {
    if (q < rank) running += tot_x[q];
    all += tot_x[q];
  }
  for (int b0 = lo; b0 < hi; b0 += blockDim.x) {
    const int b = b0 + tid;
    const uint32_t k = b < hi ? kept[rc + sp.blk0 + b] : 0u;
    int total;
    const int ex = pg_block_exscan(__popc(k), scan_smem, &total);
    uint32_t bits = k;
    int64_t dst = sp.base + running + ex;
    while (bits) {
      const int l = __ffs(bits) - 1;
      bits &= bits - 1;
      const SBox* sb = ws.sbox + (sp.blk0 + b) * 32 + l;
      PG_DEV_ASSERT(dst >= sp.base && dst < sp.base + sp.m);
      ws.kscore[dst] = sb->score;
      ws.kpos[dst] = (int32_t)sb->k;
      ++dst;
    }
    running += total;
  }
  if (rank == 0 && tid == 0) {
    n_kept[p] = all;
    atomicMax((unsigned long long*)&ws.stats[2], (unsigned long long)rounds);
  }
}

// ---- F: emit -------------------------------------------------------------------------------
constexpr int EMIT_SMEM_ELEMS = 8192;

// ascending u64 key == descending score (positive and negative doubles).  -0.0 is folded onto +0.0 first: the
// suppression ranking (prefilter_or), like Python's max()/index of 3_combine_grids.py:112, compares VALUES, so
// a -0.0 and a +0.0 score tie and the earlier pooled position wins — the emitted pick order must agree with it.
__device__ __forceinline__ unsigned long long score_desc_key(double s) {
  if (s == 0.0) s = 0.0;
  unsigned long long b = (unsigned long long)__double_as_longlong(s);
  b = (b >> 63) ? ~b : (b | 0x8000000000000000ull);  // ascending-orderable
  return ~b;
}

// Normalised bitonic network (every comparator ascending), so virtual +inf padding above K never moves.
template <typename KeyT>
__device__ __forceinline__ void bitonic_sort_pairs(KeyT* keys, int32_t* idx, int K, int n2) {
  for (int k = 2; k <= n2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < (n2 >> 1); t += blockDim.x) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int pr = (j == (k >> 1)) ? (i ^ (k - 1)) : (i ^ j);
        if (pr < K) {
          const KeyT a = keys[i], b = keys[pr];
          const int32_t ia = idx[i], ib = idx[pr];
          if (b < a || (b == a && ib < ia)) {
            keys[i] = b; keys[pr] = a; idx[i] = ib; idx[pr] = ia;
          }
        }
      }
      __syncthreads();
    }
  }
}

// ---- the same network on a shared-memory chunk, eight elements to a thread -------------------------------------
// Three consecutive steps (partner distances J, J/2, J/4) touch, for one element, only the elements that differ from
// it in those three index bits — eight elements that one thread holds in registers, so the chunk makes one trip
// through shared memory per THREE steps instead of one per step (33 trips instead of 91 for 8192 elements).  Slot
// r = (bit J, bit J/2, bit J/4) of the element's index.  The first step of a merge level k = 2J pairs i with
// i ^ (k - 1): slot r with slot r ^ 7, and the thread's elements with bit J set have their low bits flipped.
// Slots at or above the live count hold +inf keys (EMIT_PAD_KEY / EMIT_PAD_IDX) and never move below a real one.
// Storage is padded by one element in eight so that a thread's eight consecutive elements (J = 4) fall into
// different banks: element i lives at i + (i >> 3).
constexpr unsigned long long EMIT_PAD_KEY = ~0ull;
constexpr int32_t EMIT_PAD_IDX = 0x7fffffff;
constexpr int EMIT_SMEM_SLOTS = EMIT_SMEM_ELEMS + (EMIT_SMEM_ELEMS >> 3);
__device__ __forceinline__ int emit_slot(int i) { return i + (i >> 3); }

__device__ __forceinline__ void emit_cswap(unsigned long long& ka, int32_t& ia, unsigned long long& kb, int32_t& ib) {
  if (kb < ka || (kb == ka && ib < ia)) {
    const unsigned long long tk = ka; ka = kb; kb = tk;
    const int32_t ti = ia; ia = ib; ib = ti;
  }
}

// steps J, J/2, ... (nsteps <= 3 of them) for group g; flip: the first is the first step of merge level 2J
__device__ __forceinline__ void emit_pass8(unsigned long long* keys, int32_t* idx, int g, int J, int nsteps, bool flip) {
  const int q = J >> 2, lj = 31 - __clz(J);
  const int base = ((g >> (lj - 2)) << (lj + 1)) | (g & (q - 1));
  unsigned long long k[8];
  int32_t x[8];
  int at[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    int e = base | ((r >> 2) * J) | (((r >> 1) & 1) * (J >> 1)) | ((r & 1) * q);
    if (flip && (r >> 2)) e ^= q - 1;
    at[r] = emit_slot(e);
    PG_DEV_ASSERT(e >= 0 && at[r] < EMIT_SMEM_SLOTS);
    k[r] = keys[at[r]];
    x[r] = idx[at[r]];
  }
  if (flip) {
#pragma unroll
    for (int r = 0; r < 4; ++r) emit_cswap(k[r], x[r], k[r ^ 7], x[r ^ 7]);
  } else {
#pragma unroll
    for (int r = 0; r < 4; ++r) emit_cswap(k[r], x[r], k[r | 4], x[r | 4]);
  }
  if (nsteps >= 2) {
#pragma unroll
    for (int r = 0; r < 8; ++r) if (!(r & 2)) emit_cswap(k[r], x[r], k[r | 2], x[r | 2]);
  }
  if (nsteps >= 3) {
#pragma unroll
    for (int r = 0; r < 8; r += 2) emit_cswap(k[r], x[r], k[r | 1], x[r | 1]);
  }
#pragma unroll
  for (int r = 0; r < 8; ++r) { keys[at[r]] = k[r]; idx[at[r]] = x[r]; }
}

// merge steps J_top, J_top/2, ..., 1 of one level over n elements (n a power of two >= 16, J_top >= 8 or == 4):
// a first pass of 1-3 steps so that three-step passes finish exactly at distance 1
__device__ __forceinline__ void emit_merge_steps(unsigned long long* keys, int32_t* idx, int n, int J_top, bool flip) {
  const int L = 32 - __clz(J_top);  // steps to go
  int J = J_top, s = ((L - 1) % 3) + 1;
  bool fl = flip;
  while (J >= 1) {
    for (int g = threadIdx.x; g < (n >> 3); g += blockDim.x) emit_pass8(keys, idx, g, J, s, fl);
    __syncthreads();
    J >>= s;
    s = 3;
    fl = false;
  }
}

// whole sort of n elements in padded shared storage (n a power of two, 16 <= n <= EMIT_SMEM_ELEMS)
__device__ __forceinline__ void emit_sort_chunk(unsigned long long* keys, int32_t* idx, int n) {
  // levels 2, 4, 8 on eight consecutive elements
  for (int g = threadIdx.x; g < (n >> 3); g += blockDim.x) {
    unsigned long long k[8];
    int32_t x[8];
    const int s0 = emit_slot(8 * g);
#pragma unroll
    for (int r = 0; r < 8; ++r) { k[r] = keys[s0 + r]; x[r] = idx[s0 + r]; }
#pragma unroll
    for (int r = 0; r < 8; r += 2) emit_cswap(k[r], x[r], k[r | 1], x[r | 1]);          // k = 2
#pragma unroll
    for (int r = 0; r < 8; ++r) if (!(r & 2)) emit_cswap(k[r], x[r], k[r ^ 3], x[r ^ 3]);  // k = 4: flip, then 1
#pragma unroll
    for (int r = 0; r < 8; r += 2) emit_cswap(k[r], x[r], k[r | 1], x[r | 1]);
#pragma unroll
    for (int r = 0; r < 4; ++r) emit_cswap(k[r], x[r], k[r ^ 7], x[r ^ 7]);              // k = 8: flip, then 2, 1
#pragma unroll
    for (int r = 0; r < 8; ++r) if (!(r & 2)) emit_cswap(k[r], x[r], k[r | 2], x[r | 2]);
#pragma unroll
    for (int r = 0; r < 8; r += 2) emit_cswap(k[r], x[r], k[r | 1], x[r | 1]);
#pragma unroll
    for (int r = 0; r < 8; ++r) { keys[s0 + r] = k[r]; idx[s0 + r] = x[r]; }
  }
  __syncthreads();
  for (int k2 = 16; k2 <= n; k2 <<= 1) emit_merge_steps(keys, idx, n, k2 >> 1, true);
}

// `count` live elements of a chunk of `n` into padded shared storage, +inf behind them.  as_scores: the source
// holds scores still (they become sortable keys on the way in); otherwise keys a previous stage wrote (L2 loads).
__device__ __forceinline__ void emit_load_chunk(unsigned long long* keys, int32_t* idx, const double* src, const int32_t* src_idx,
                                                int count, int n, bool as_scores) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    unsigned long long k = EMIT_PAD_KEY;
    int32_t x = EMIT_PAD_IDX;
    if (i < count) {
      if (as_scores) { k = score_desc_key(src[i]); x = src_idx[i]; }
      else { k = __ldcg(reinterpret_cast<const unsigned long long*>(src) + i); x = __ldcg(src_idx + i); }
    }
    keys[emit_slot(i)] = k;
    idx[emit_slot(i)] = x;
  }
  __syncthreads();
}
// after emit_load_chunk: the whole network on n elements (tiny n: the one-step-per-trip form; slots == indices below 8)
__device__ __forceinline__ void emit_sort_loaded(unsigned long long* keys, int32_t* idx, int count, int n) {
  if (n >= 16) emit_sort_chunk(keys, idx, n);
  else bitonic_sort_pairs(keys, idx, count, n);
}

__global__ void __launch_bounds__(1024) nms_emit_kernel(const int32_t* __restrict__ sel_idx,
                                                        const int64_t* __restrict__ page_off, NmsWs ws,
                                                        const int32_t* __restrict__ n_kept,
                                                        int32_t* __restrict__ kept_idx) {
  extern __shared__ __align__(16) unsigned char emit_smem[];
  const int p = blockIdx.x, tid = threadIdx.x;
  const int K = n_kept[p];
  if (K <= 0) return;
  const int64_t base = page_off[p];
  int n2 = 1;
  while (n2 < K) n2 <<= 1;
  if (n2 <= EMIT_SMEM_ELEMS) {
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(emit_smem);
    int32_t* idx = reinterpret_cast<int32_t*>(emit_smem + (size_t)EMIT_SMEM_SLOTS * 8);
    emit_load_chunk(keys, idx, ws.kscore + base, ws.kpos + base, K, n2, true);
    emit_sort_loaded(keys, idx, K, n2);
    for (int i = tid; i < K; i += blockDim.x) {
      const int k = idx[emit_slot(i)];
      kept_idx[base + i] = sel_idx ? sel_idx[base + k] : (int32_t)(base + k);
    }
  } else {
    // large pages: the same network straight on the (L2-resident) workspace arrays
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(ws.kscore + base);
    int32_t* idx = ws.kpos + base;
    for (int i = tid; i < K; i += blockDim.x) keys[i] = score_desc_key(__longlong_as_double((long long)keys[i]));
    __syncthreads();
    bitonic_sort_pairs(keys, idx, K, n2);
    for (int i = tid; i < K; i += blockDim.x) {
      const int k = idx[i];
      kept_idx[base + i] = sel_idx ? sel_idx[base + k] : (int32_t)(base + k);
    }
  }
}

// ---- F (large pages): emit on one thread-block cluster per page -------------------------------
// The same normalised bitonic network on K > 8192 survivors, cut at chunks of EMIT_SMEM_ELEMS:
// every step whose partner distance stays inside a chunk runs in shared memory (each CTA owns the
// chunks rank, rank + csize, ...), the few steps that cross chunks run on the L2-resident arrays with
// all CTAs of the cluster sharing the comparators; a cluster barrier separates the stages.  Elements at
// or above K are virtual +inf: a comparator (i, pr) has i < pr and is skipped when pr >= K.
__device__ __forceinline__ void emit_cswap_global(unsigned long long* keys, int32_t* idx, int i, int pr) {
  const unsigned long long a = __ldcg(keys + i), b = __ldcg(keys + pr);
  const int32_t ia = __ldcg(idx + i), ib = __ldcg(idx + pr);
  if (b < a || (b == a && ib < ia)) {
    __stcg(keys + i, b); __stcg(keys + pr, a); __stcg(idx + i, ib); __stcg(idx + pr, ia);
  }
}

__global__ void __launch_bounds__(1024) nms_emit_cluster_kernel(const int32_t* __restrict__ sel_idx,
                                                                const int64_t* __restrict__ page_off, NmsWs ws,
                                                                const int32_t* __restrict__ n_kept,
                                                                int32_t* __restrict__ kept_idx) {
  extern __shared__ __align__(16) unsigned char emit_smem[];
  cg::cluster_group cluster = cg::this_cluster();
  const int csize = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
  const int p = blockIdx.x / csize, tid = threadIdx.x;
  const int K = n_kept[p];  // the same for every CTA of the cluster
  if (K <= 0) return;
  const int64_t base = page_off[p];
  int n2 = 1;
  while (n2 < K) n2 <<= 1;
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(emit_smem);
  int32_t* idx = reinterpret_cast<int32_t*>(emit_smem + (size_t)EMIT_SMEM_SLOTS * 8);
  unsigned long long* gkeys = reinterpret_cast<unsigned long long*>(ws.kscore + base);
  int32_t* gidx = ws.kpos + base;
  if (n2 <= EMIT_SMEM_ELEMS) {  // few survivors on this page: one CTA, no cluster barrier on this path for anyone
    if (rank != 0) return;
    emit_load_chunk(keys, idx, ws.kscore + base, ws.kpos + base, K, n2, true);
    emit_sort_loaded(keys, idx, K, n2);
    for (int i = tid; i < K; i += blockDim.x) {
      const int k = idx[emit_slot(i)];
      kept_idx[base + i] = sel_idx ? sel_idx[base + k] : (int32_t)(base + k);
    }
    return;
  }
  int CH = EMIT_SMEM_ELEMS;  // chunk of the network that one CTA sorts in shared memory: every CTA of the cluster gets one
  while (CH > 2048 && n2 / CH < csize) CH >>= 1;
  const int nch = n2 / CH;
  // stage 0: every chunk sorted on its own (k = 2 .. CH), scores turned into sortable keys on the way in
  for (int c = rank; c < nch; c += csize) {
    const int cb = c * CH;
    if (cb >= K) break;
    const int kc = min(CH, K - cb);
    emit_load_chunk(keys, idx, ws.kscore + base + cb, gidx + cb, kc, CH, true);
    emit_sort_chunk(keys, idx, CH);
    for (int i = tid; i < kc; i += blockDim.x) { __stcg(gkeys + cb + i, keys[emit_slot(i)]); __stcg(gidx + cb + i, idx[emit_slot(i)]); }
    __syncthreads();
  }
  __threadfence();
  cluster.sync();
  for (int k = 2 * CH; k <= n2; k <<= 1) {
    // steps that cross chunks: all CTAs of the cluster, straight on the workspace arrays
    for (int j = k >> 1; j >= CH; j >>= 1) {
      for (int t = rank * (int)blockDim.x + tid; t < (n2 >> 1); t += csize * (int)blockDim.x) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int pr = (j == (k >> 1)) ? (i ^ (k - 1)) : (i ^ j);
        if (pr < K) emit_cswap_global(gkeys, gidx, i, pr);
      }
      __threadfence();
      cluster.sync();
    }
    // the remaining steps j = CH/2 .. 1 stay inside a chunk
    const bool last = (k == n2);
    for (int c = rank; c < nch; c += csize) {
      const int cb = c * CH;
      if (cb >= K) break;
      const int kc = min(CH, K - cb);
      emit_load_chunk(keys, idx, reinterpret_cast<const double*>(gkeys + cb), gidx + cb, kc, CH, false);
      emit_merge_steps(keys, idx, CH, CH >> 1, false);
      if (last) {
        for (int i = tid; i < kc; i += blockDim.x) {
          const int kk = idx[emit_slot(i)];
          kept_idx[base + cb + i] = sel_idx ? sel_idx[base + kk] : (int32_t)(base + kk);
        }
      } else {
        for (int i = tid; i < kc; i += blockDim.x) { __stcg(gkeys + cb + i, keys[emit_slot(i)]); __stcg(gidx + cb + i, idx[emit_slot(i)]); }
      }
      __syncthreads();
    }
    if (!last) {
      __threadfence();
      cluster.sync();
    }
  }
}

// Cluster width for the resolve/emit pair of a launch: 1 = one CTA per page (the default kernels).
// Clusters pay off when a page is large and the launch has fewer pages than SMs.
static int nms_cluster_width(int32_t n_pages, int32_t max_boxes_per_page, int sms, int emit_smem) {
  int min_boxes = 32768;
  if (const char* e = getenv("PG_NMS_CLUSTER_MIN_BOXES")) min_boxes = atoi(e);  // test / tuning knob; <= 0 disables
  if (min_boxes <= 0 || max_boxes_per_page < min_boxes) return 1;
  int c = NMS_CLUSTER_MAX;
  if (const char* e = getenv("PG_NMS_CLUSTER_MAX")) c = std::max(1, std::min(NMS_CLUSTER_MAX, atoi(e)));  // tuning knob
  while (c & (c - 1)) c &= c - 1;
  while (c > 1 && (int64_t)n_pages * c > sms) c >>= 1;
  // the device must be able to co-schedule such clusters (1024 threads and the 96 KB emit buffer per CTA); above the
  // portable width of 8 every page's cluster must be resident at once (a cluster lives inside one GPC), or the
  // pages would queue behind each other and the wider cluster would lose what it gains
  static int active[NMS_CLUSTER_MAX + 1] = {0};  // 0 unknown, else 1 + clusters the device keeps resident (min over the three kernels)
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(nms_bin_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaFuncSetAttribute(nms_resolve_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaFuncSetAttribute(nms_emit_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaGetLastError();
    attr_set = true;
  }
  for (; c > 1; c >>= 1) {
    if (active[c] == 0) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3((unsigned)(c * (sms / c)));
      cfg.blockDim = dim3(1024);
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = (unsigned)c; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at;
      cfg.numAttrs = 1;
      int n_min = 1 << 30;
      cfg.dynamicSmemBytes = (size_t)emit_smem;
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, nms_emit_cluster_kernel, &cfg) != cudaSuccess) { cudaGetLastError(); n = 0; }
      n_min = std::min(n_min, n);
      cfg.dynamicSmemBytes = 0;
      if (cudaOccupancyMaxActiveClusters(&n, nms_resolve_cluster_kernel, &cfg) != cudaSuccess) { cudaGetLastError(); n = 0; }
      n_min = std::min(n_min, n);
      if (cudaOccupancyMaxActiveClusters(&n, nms_bin_cluster_kernel, &cfg) != cudaSuccess) { cudaGetLastError(); n = 0; }
      n_min = std::min(n_min, n);
      active[c] = 1 + n_min;
    }
    const int resident = active[c] - 1;
    if (resident > 0 && (c <= 8 || resident >= n_pages)) return c;
  }
  return 1;
}

template <typename... KArgs, typename... Args>
static cudaError_t launch_cluster(void (*kernel)(KArgs...), int n_pages, int width, size_t smem, cudaStream_t s,
                                  Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(n_pages * width));
  cfg.blockDim = dim3(1024);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)width; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

extern "C" int pg_nms_merge(const double* boxes, const double* scores, const double* classes,
                            const int32_t* sel_idx, const int64_t* page_off, const int32_t* n_sel,
                            int32_t n_pages, int64_t n_boxes, int32_t max_boxes_per_page, double iou_threshold,
                            int32_t* kept_idx, int32_t* n_kept, void* workspace, size_t workspace_bytes,
                            void* stream) {
  return pg_nms_merge_ex(boxes, scores, classes, sel_idx, page_off, n_sel, n_pages, n_boxes, max_boxes_per_page,
                         iou_threshold, 0, kept_idx, n_kept, workspace, workspace_bytes, stream);
}

extern "C" int pg_nms_merge_ex(const double* boxes, const double* scores, const double* classes,
                               const int32_t* sel_idx, const int64_t* page_off, const int32_t* n_sel,
                               int32_t n_pages, int64_t n_boxes, int32_t max_boxes_per_page, double iou_threshold,
                               int32_t mode, int32_t* kept_idx, int32_t* n_kept, void* workspace,
                               size_t workspace_bytes, void* stream) {
  PG_REQUIRE(n_pages >= 0 && n_boxes >= 0, "sizes");
  PG_REQUIRE((mode & ~(PG_NMS_CLASS_AGNOSTIC | PG_NMS_FP32)) == 0, "mode");
  if (n_pages == 0) return PG_OK;
  PG_REQUIRE(boxes && scores && (classes || (mode & PG_NMS_CLASS_AGNOSTIC)) && page_off && kept_idx && n_kept && workspace,
             "null device pointer");
  PG_REQUIRE(n_boxes < (1ll << 31), "n_boxes must fit int32");
  PG_REQUIRE(((uintptr_t)workspace & 255) == 0, "workspace must be 256-byte aligned");
  // recover pairs_per_block from the size the caller allocated
  NmsWs ws;
  const size_t fixed = nms_layout(n_boxes, n_pages, 1, nullptr, &ws);
  const size_t per_pair_block = (size_t)ws.nb_cap * (4 + 128);
  if (workspace_bytes < fixed) {
    pg_set_error("workspace too small: %zu < %zu", workspace_bytes, fixed);
    return PG_ERR_WORKSPACE;
  }
  int64_t ppb = 1 + (int64_t)((workspace_bytes - fixed) / (per_pair_block + 512));
  while (ppb > 1 && nms_layout(n_boxes, n_pages, (int32_t)ppb, nullptr, nullptr) > workspace_bytes) --ppb;
  nms_layout(n_boxes, n_pages, (int32_t)ppb, (uint8_t*)workspace, &ws);
  ws.mode = mode;
  ws.fine_min = NMS_FINE_MIN_PER_CELL;
  if (const char* e = getenv("PG_NMS_FINE_MIN")) ws.fine_min = atoi(e);  // tuning knob; a huge value disables the third pass
  cudaStream_t s = (cudaStream_t)stream;
  const int emit_smem = EMIT_SMEM_SLOTS * 12;
  PG_CUDA_TRY(cudaFuncSetAttribute(nms_emit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, emit_smem));
  PG_CUDA_TRY(cudaFuncSetAttribute(nms_emit_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, emit_smem));
  const int all_pairs = !(iou_threshold >= 0.0);  // thr < 0: disjoint boxes (IoU 0) suppress too
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // per-page kernels (bin, resolve, emit): one CTA per page, or one cluster per page for few large pages
  const int width = nms_cluster_width(n_pages, max_boxes_per_page, sms, emit_smem);
  if (getenv("PG_NMS_DEBUG")) fprintf(stderr, "pg_nms_merge: %d pages, cluster width %d\n", (int)n_pages, width);
  if (width > 1) {
    PG_CUDA_TRY(launch_cluster(nms_bin_cluster_kernel, n_pages, width, 0, s, boxes, scores, classes, sel_idx, page_off,
                               n_sel, (int)n_pages, ws));
  } else {
    nms_bin_kernel<<<n_pages, 1024, 0, s>>>(boxes, scores, classes, sel_idx, page_off, n_sel, n_pages, ws);
    PG_LAUNCH_CHECK();
  }
  const unsigned cand_grid = (unsigned)((ws.nb_cap + 7) / 8);
  nms_cand_kernel<<<cand_grid, 256, 0, s>>>(page_off, n_sel, ws, all_pairs);
  PG_LAUNCH_CHECK();
  const int64_t mask_want = (ws.ent_cap + 8 * MASK_UNIT - 1) / (8 * MASK_UNIT);
  int mask_per_sm = 3;
  // 3 resident CTAs per SM (80 registers) for ordinary pages; 4 (64 registers) on the crowded pages that also take
  // the cluster kernels, where the mask kernel is most of the call (cfg4: 0.844 -> 0.822 ms)
  int mask_occ = width > 1 ? 4 : 3;
  if (const char* e = getenv("PG_NMS_MASK_OCC")) mask_occ = atoi(e);  // tuning knob
  void (*mask_kernel)(NmsWs, double) = mask_occ == 4 ? nms_mask_kernel<4> : nms_mask_kernel<3>;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&mask_per_sm, mask_kernel, 256, 0);
  const int64_t mask_slots = (int64_t)sms * (mask_per_sm > 0 ? mask_per_sm : 1);
  const unsigned mask_grid = (unsigned)(mask_want < mask_slots ? (mask_want < 1 ? 1 : mask_want) : mask_slots);
  mask_kernel<<<mask_grid, 256, 0, s>>>(ws, iou_threshold);
  PG_LAUNCH_CHECK();
  if (width > 1) {
    PG_CUDA_TRY(launch_cluster(nms_resolve_cluster_kernel, n_pages, width, 0, s, page_off, n_sel, ws, n_kept));
    PG_CUDA_TRY(launch_cluster(nms_emit_cluster_kernel, n_pages, width, (size_t)emit_smem, s, sel_idx, page_off, ws,
                               (const int32_t*)n_kept, kept_idx));
    return PG_OK;
  }
  nms_resolve_kernel<<<n_pages, page_kernel_threads(1024), 0, s>>>(page_off, n_sel, ws, n_kept);
  PG_LAUNCH_CHECK();
  nms_emit_kernel<<<n_pages, page_kernel_threads(1024), emit_smem, s>>>(sel_idx, page_off, ws, n_kept, kept_idx);
  PG_LAUNCH_CHECK();
  return PG_OK;
}

extern "C" int pg_nms_stats(const void* workspace, int64_t stats[4]) {
  PG_REQUIRE(workspace && stats, "null pointer");
  int64_t h[8];
  PG_CUDA_TRY(cudaMemcpy(h, workspace, sizeof(h), cudaMemcpyDeviceToHost));
  stats[0] = h[ST_STATUS];
  stats[1] = h[ST_ENT_TOTAL];
  stats[2] = h[ST_ROUNDS];
  stats[3] = h[ST_TESTS];
  return PG_OK;
}

// =============================================================================================
// K4 — plain_text width bins + median
//   reference: bin_widths 4_extract_median_widths.py:49-80 (sequential leader binning, joins the
//   smallest-key bin within the margin), calculate_median_width :82-101 (np.median), width
//   extraction :135-141.
// One CTA per page.  All warps first gather the plain_text widths in order (independent loads);
// warp 0 then bins them: 32 widths are matched against the sorted bin keys at once (binary search
// on the reference's own fabs(w-key) <= margin predicate); only a width that founds a new bin
// serialises, and lanes behind it re-check just that new key.
// =============================================================================================
__global__ void class_flags_kernel(const double* __restrict__ classes, int64_t n, double plain_id, double title_id,
                                   uint8_t* __restrict__ flags) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double c = classes[i];
    flags[i] = (uint8_t)((c == plain_id ? PG_FLAG_PLAIN_TEXT : 0u) | (c == title_id ? PG_FLAG_TITLE : 0u));
  }
}

extern "C" int pg_class_flags(const double* classes, int64_t n, double plain_text_id, double title_id,
                              uint8_t* flags, void* stream) {
  PG_REQUIRE(n >= 0, "n");
  if (n == 0) return PG_OK;
  PG_REQUIRE(classes && flags, "null device pointer");
  const unsigned grid = (unsigned)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
  class_flags_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(classes, n, plain_text_id, title_id, flags);
  PG_LAUNCH_CHECK();
  return PG_OK;
}

__device__ __forceinline__ int bins_find(const double* keys, int n, double w, double margin) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    const double kk = keys[mid];
    if (kk >= w || (w - kk) <= margin) hi = mid; else lo = mid + 1;
  }
  if (lo < n && fabs(w - keys[lo]) <= margin) return lo;  // 4_extract_median_widths.py:71
  return -1;
}

constexpr int MED_THREADS = 256;
__global__ void __launch_bounds__(MED_THREADS) width_median_kernel(
    const double* __restrict__ boxes, const uint8_t* __restrict__ flags, const int32_t* __restrict__ sel_idx,
    const int64_t* __restrict__ page_off, const int32_t* __restrict__ n_sel, const int32_t* __restrict__ page_wh,
    double margin_pct, double* __restrict__ median, int32_t* __restrict__ n_bins, double* ws_keys, int32_t* ws_counts,
    int64_t n_total, uint32_t* width_hist) {
  __shared__ int scan_smem[34];
  const int p = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
  const int64_t base = page_off[p];
  const int m = n_sel ? n_sel[p] : (int)(page_off[p + 1] - base);
  double* keys = ws_keys + base;
  double* wlist = ws_keys + n_total + base;
  int* counts = ws_counts + base;
  const double margin = (double)page_wh[2 * p] * (margin_pct / 100.0);  // :64

  // ---- gather the plain_text widths in pooled order (:135-141) ----
  int nw = 0;
  for (int c = 0; c < m; c += MED_THREADS) {
    const int k = c + tid;
    int act = 0;
    double w = 0.0;
    if (k < m) {
      const int64_t gi = sel_idx ? (int64_t)sel_idx[base + k] : base + k;
      if (flags[gi] & PG_FLAG_PLAIN_TEXT) {
        act = 1;
        w = boxes[4 * gi + 2] - boxes[4 * gi];  // :139
      }
    }
    if (width_hist) {  // corpus histogram (K6): one atomic per distinct bin per warp (widths cluster heavily)
      const int hb = act ? (w >= 0.0 ? (w < (double)(PG_WIDTH_HIST_BINS - 1) ? (int)w : PG_WIDTH_HIST_BINS - 1) : 0)
                         : -1 - lane;
      const unsigned grp = __match_any_sync(0xffffffffu, hb);
      if (act && lane == __ffs(grp) - 1) atomicAdd(&width_hist[hb], (unsigned)__popc(grp));
    }
    int total;
    const int ex = pg_block_exscan(act, scan_smem, &total);
    if (act) wlist[nw + ex] = w;
    nw += total;
  }
  __syncthreads();
  if (tid >= 32) return;

  // ---- sequential-equivalent leader binning, 32 widths at a time ----
  int nb = 0;
  for (int c = 0; c < nw; c += 32) {
    const bool act = c + lane < nw;
    const double w = act ? wlist[c + lane] : 0.0;
    unsigned pend = __ballot_sync(0xffffffffu, act);
    int mt = act ? bins_find(keys, nb, w, margin) : -1;
    while (pend) {
      const bool mine = (pend >> lane) & 1u;
      const unsigned nomatch = __ballot_sync(0xffffffffu, mine && mt < 0);
      const int first = nomatch ? __ffs(nomatch) - 1 : 32;
      const unsigned fin = first == 32 ? pend : (pend & ((1u << first) - 1u));
      if ((fin >> lane) & 1u) atomicAdd(&counts[mt], 1);  // :72
      pend &= ~fin;
      __syncwarp();
      if (first == 32) break;
      // lane `first` founds a bin keyed by its own width (:77-78)
      const double wf = __shfl_sync(0xffffffffu, w, first);
      int ins = 0;
      for (int b = lane; b < nb; b += 32) ins += keys[b] < wf ? 1 : 0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) ins += __shfl_xor_sync(0xffffffffu, ins, o);
      const bool exists = ins < nb && keys[ins] == wf;  // only reachable with a negative margin
      if (exists) {
        if (lane == 0) counts[ins] = 1;  // dict assignment resets the count
      } else {
        for (int hi = nb; hi > ins; hi -= 32) {
          const int idx = hi - 1 - lane;
          double tk = 0.0;
          int tc = 0;
          if (idx >= ins) { tk = keys[idx]; tc = counts[idx]; }
          __syncwarp();
          if (idx >= ins) { keys[idx + 1] = tk; counts[idx + 1] = tc; }
          __syncwarp();
        }
        if (lane == 0) { keys[ins] = wf; counts[ins] = 1; }
        ++nb;
      }
      __syncwarp();
      pend &= ~(1u << first);
      if (mine && lane != first) {
        if (!exists && mt >= ins) mt += 1;
        if (fabs(w - wf) <= margin && (mt < 0 || ins < mt)) mt = ins;
      }
    }
  }
  // ---- median over keys repeated by count; keys are already ascending ----
  int total = 0;
  for (int b = lane; b < nb; b += 32) total += counts[b];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
  double med = 0.0;
  if (total > 0) {
    const int r_lo = (total - 1) >> 1, r_hi = total >> 1;
    double v_lo = 0.0, v_hi = 0.0;
    int carry = 0;
    for (int b0 = 0; b0 < nb; b0 += 32) {
      const int b = b0 + lane;
      const int cnt = b < nb ? counts[b] : 0;
      int inc = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
      }
      const int lo_ex = carry + inc - cnt, hi_ex = carry + inc;
      const bool has_lo = cnt > 0 && r_lo >= lo_ex && r_lo < hi_ex;
      const bool has_hi = cnt > 0 && r_hi >= lo_ex && r_hi < hi_ex;
      const unsigned bl = __ballot_sync(0xffffffffu, has_lo), bh = __ballot_sync(0xffffffffu, has_hi);
      const double kv = b < nb ? keys[b] : 0.0;
      if (bl) v_lo = __shfl_sync(0xffffffffu, kv, __ffs(bl) - 1);
      if (bh) v_hi = __shfl_sync(0xffffffffu, kv, __ffs(bh) - 1);
      carry = __shfl_sync(0xffffffffu, hi_ex, 31);
    }
    med = (total & 1) ? v_lo : (v_lo + v_hi) / 2.0;  // np.median
  }
  if (lane == 0) { median[p] = med; n_bins[p] = nb; }
}

extern "C" int pg_width_median(const double* boxes, const uint8_t* flags, const int32_t* sel_idx,
                               const int64_t* page_off, const int32_t* n_sel, int32_t n_pages, int64_t n_boxes,
                               const int32_t* page_wh, double min_margin_percent, double* median,
                               int32_t* n_bins, double* ws_keys, int32_t* ws_counts, uint32_t* width_hist,
                               void* stream) {
  PG_REQUIRE(n_pages >= 0 && n_boxes >= 0, "sizes");
  if (n_pages == 0) return PG_OK;
  PG_REQUIRE(boxes && flags && page_off && page_wh && median && n_bins && ws_keys && ws_counts, "null device pointer");
  width_median_kernel<<<n_pages, MED_THREADS, 0, (cudaStream_t)stream>>>(
      boxes, flags, sel_idx, page_off, n_sel, page_wh, min_margin_percent, median, n_bins, ws_keys, ws_counts, n_boxes,
      width_hist);
  PG_LAUNCH_CHECK();
  return PG_OK;
}

// =============================================================================================
// K5 — column centres
//   reference: find_column_centers 5_detect_column_centers.py:91-224 (+ scipy find_peaks step
//   order: local maxima -> height -> distance -> prominence; np.convolve 'same').
// Two kernels.  (a) density: each warp owns 32 consecutive bins of one page and walks the
// page's accepted boxes in pooled order (block-scan compaction, then ballot over the boxes that
// overlap the warp's bins), so every bin receives its `density[bin] += w` contributions in the
// reference's order and the map is bit-identical to numpy's; the per-bin divide is the exact
// reciprocal+FMA form of pg_math.h.  (b) peaks: one CTA per page — left-to-right smoothing sum,
// maximum, plateau-aware local maxima, scipy's distance / prominence filters, valley walk.
// =============================================================================================
constexpr int DEN_THREADS = 256;  // 8 warps x 32 bins per CTA
constexpr int COL_THREADS = 1024;
constexpr int COL_BPT = 2;        // peaks kernel: bins per thread -> 2048 bins (>= the 2000-bin bound)
constexpr int COL_MAX_PEAKS = 1024;

struct ColGeom {
  int res, nbins, win;
  bool ok;       // page passes the process_page guards
  bool in_range; // shape inside the kernel limits
};

__device__ __forceinline__ ColGeom col_geom(int W, int Hh, double med, int max_bins, int max_window) {
  ColGeom g;
  g.ok = (med > 0.0) && W > 0 && Hh > 0;  // 5_detect_column_centers.py:361-364, 381-383
  g.res = max(1, W / 1000);                                 // :120
  g.nbins = W / g.res + 1;                                  // :121
  g.win = 5;
  if (g.ok) {
    const double wv = med / (4.0 * (double)g.res);         // :147
    g.win = wv >= 5.0 ? (wv < 1.0e9 ? (int)wv : 1000000000) : 5;
    if ((g.win & 1) == 0) g.win += 1;
  }
  g.in_range = g.nbins <= max_bins && g.nbins <= COL_THREADS * COL_BPT && g.win <= max_window && g.win <= g.nbins;
  return g;
}

struct __align__(16) DenEntry {
  int left, right, center, pad;
  double half, inv_half;
};

// (a1) prep: one CTA per page turns the page's boxes into the ordered list of accepted bin spans
__global__ void __launch_bounds__(DEN_THREADS) column_prep_kernel(
    const double* __restrict__ boxes, const uint8_t* __restrict__ flags, const double* __restrict__ scores,
    const int32_t* __restrict__ sel_idx, const int64_t* __restrict__ page_off, const int32_t* __restrict__ n_sel,
    const int32_t* __restrict__ page_wh, const double* __restrict__ median, int max_window, double min_conf,
    int max_bins, DenEntry* __restrict__ list, int32_t* __restrict__ list_n) {
  __shared__ int scan_smem[34];
  const int p = blockIdx.x, tid = threadIdx.x;
  const int W = page_wh[2 * p], Hh = page_wh[2 * p + 1];
  const double med = median[p];
  const ColGeom g = col_geom(W, Hh, med, max_bins, max_window);
  if (!g.ok || !g.in_range) {
    if (tid == 0) list_n[p] = 0;
    return;
  }
  const int64_t base = page_off[p];
  const int m = n_sel ? n_sel[p] : (int)(page_off[p + 1] - base);
  const double lo_w = 0.33 * med, hi_w = 2.0 * med;  // :131
  int running = 0;
  for (int c0 = 0; c0 < m; c0 += DEN_THREADS) {
    const int k = c0 + tid;
    int acc = 0;
    DenEntry ent;
    ent.left = 1; ent.right = -1; ent.center = 0; ent.pad = 0; ent.half = 1.0; ent.inv_half = 1.0;
    if (k < m) {
      const int64_t gi = sel_idx ? (int64_t)sel_idx[base + k] : base + k;
      if ((flags[gi] & (PG_FLAG_PLAIN_TEXT | PG_FLAG_TITLE)) && scores[gi] >= min_conf) {  // :110-112
        const int x1 = (int)boxes[4 * gi], x2 = (int)boxes[4 * gi + 2];                    // :127 int() truncation
        const int bw = x2 - x1;
        if (lo_w <= (double)bw && (double)bw <= hi_w) {
          ent.left = max(0, pg_floordiv(x1, g.res));                 // :133
          ent.right = min(g.nbins - 1, pg_floordiv(x2, g.res));      // :134
          ent.center = pg_floordiv(x1 + x2, 2 * g.res);              // :137
          if (ent.left <= ent.right) {
            acc = 1;
            ent.half = pg_density_half(ent.left, ent.right);
            ent.inv_half = 1.0 / ent.half;
            // outside the exhaustively verified domain of the reciprocal form: flag for the true divide
            const int far = max(abs(ent.left - ent.center), abs(ent.right - ent.center));
            ent.pad = (ent.right - ent.left > PG_RCP_DOMAIN || far > PG_RCP_DOMAIN) ? 1 : 0;
          }
        }
      }
    }
    int total;
    const int ex = pg_block_exscan(acc, scan_smem, &total);
    if (acc) list[base + running + ex] = ent;
    running += total;
  }
  if (tid == 0) list_n[p] = running;
}

// (a2) density: each warp owns 32 bins and streams the page's span list in order
__global__ void __launch_bounds__(DEN_THREADS) column_density_kernel(
    const int64_t* __restrict__ page_off, const int32_t* __restrict__ page_wh, const double* __restrict__ median,
    int max_window, const DenEntry* __restrict__ list, const int32_t* __restrict__ list_n, double* ws_all,
    int max_bins) {
  // The eight warps of a CTA walk the same span list: it is staged through shared memory by the whole CTA,
  // 256 spans at a time, double-buffered (the next chunk's loads are in flight under this chunk's arithmetic).
  __shared__ DenEntry buf[2][DEN_THREADS];
  const int p = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const ColGeom g = col_geom(page_wh[2 * p], page_wh[2 * p + 1], median[p], max_bins, max_window);
  if (!g.ok || !g.in_range || (int)blockIdx.x * DEN_THREADS >= g.nbins) return;
  const DenEntry* lst = list + page_off[p];
  const int n = list_n[p];
  const int seg0 = ((int)blockIdx.x * (DEN_THREADS / 32) + warp) * 32;
  const int bin = seg0 + lane;
  double d = 0.0;
  DenEntry none;  // overlaps nothing (right < 0 <= seg0)
  none.left = 1; none.right = -1; none.center = 0; none.pad = 0; none.half = 1.0; none.inv_half = 1.0;
  const int n_chunks = (n + DEN_THREADS - 1) / DEN_THREADS;
  if (n_chunks > 0) buf[0][tid] = tid < n ? lst[tid] : none;
  __syncthreads();
  for (int c = 0; c < n_chunks; ++c) {
    const DenEntry* cur = buf[c & 1];
    DenEntry pre = none;
    const int nx = (c + 1) * DEN_THREADS + tid;
    if (nx < n) pre = lst[nx];
    for (int sub = 0; sub < DEN_THREADS / 32 && c * DEN_THREADS + sub * 32 < n; ++sub) {
      const DenEntry* grp = cur + sub * 32;
      unsigned bits = __ballot_sync(0xffffffffu, grp[lane].left <= seg0 + 31 && grp[lane].right >= seg0);
      // Four spans at a time: their weights are independent (ILP hides the fp64 latency), the adds stay
      // in box order (:139-144).  A bin outside a span adds +0.0, which leaves any d >= 0 bit-identical.
      while (bits) {
        DenEntry en[4];
        bool ok[4];
        int slow = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          ok[q] = bits != 0;
          const int l = ok[q] ? __ffs(bits) - 1 : 0;
          bits &= bits - 1;              // 0 stays 0
          en[q] = grp[l];                // broadcast
          slow |= en[q].pad;
        }
        double w4[4];
        if (!slow) {  // warp-uniform; straight-line code so the four chains interleave
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const bool in = ok[q] && bin >= en[q].left && bin <= en[q].right;
            const double wgt = pg_density_weight_rcp(bin, en[q].center, en[q].half, en[q].inv_half);
            w4[q] = in ? wgt : 0.0;
          }
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const bool in = ok[q] && bin >= en[q].left && bin <= en[q].right;
            w4[q] = 0.0;
            if (in) w4[q] = en[q].pad ? pg_density_weight(bin, en[q].left, en[q].right, en[q].center)
                                      : pg_density_weight_rcp(bin, en[q].center, en[q].half, en[q].inv_half);
          }
        }
        d = d + w4[0];
        d = d + w4[1];
        d = d + w4[2];
        d = d + w4[3];
      }
    }
    buf[(c + 1) & 1][tid] = pre;  // the other buffer: nobody reads it during this chunk
    __syncthreads();
  }
  if (bin < g.nbins) ws_all[(int64_t)p * 2 * max_bins + bin] = d;
}


__device__ __forceinline__ double block_max_d(double v, double* red) {
  v = warp_max_d(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    double t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : -DBL_MAX;
    t = warp_max_d(t);
    if (threadIdx.x == 0) red[32] = t;
  }
  __syncthreads();
  const double r = red[32];
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(COL_THREADS) column_peaks_kernel(
    const int32_t* __restrict__ page_wh, const double* __restrict__ median, const double* __restrict__ gauss_table,
    const int64_t* __restrict__ gauss_off, int max_window, int max_cols, int32_t* __restrict__ centers,
    double* __restrict__ widths, int32_t* __restrict__ n_cols, double* ws_all, int max_bins, uint32_t* col_hist) {
  __shared__ int scan_smem[34];
  __shared__ double red[33];
  __shared__ int pk_pos[COL_MAX_PEAKS];
  __shared__ double pk_h[COL_MAX_PEAKS];
  __shared__ unsigned char pk_keep[COL_MAX_PEAKS];
  __shared__ double sm_s[COL_THREADS * COL_BPT];

  const int p = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int W = page_wh[2 * p], Hh = page_wh[2 * p + 1];
  const double med = median[p];
  const ColGeom g = col_geom(W, Hh, med, max_bins, max_window);
  if (!g.ok || !g.in_range) {
    if (tid == 0) n_cols[p] = g.ok ? -1 : 0;
    return;
  }
  const int res = g.res, nbins = g.nbins;
  const double* dens = ws_all + (int64_t)p * 2 * max_bins;
  double* sm_out = ws_all + (int64_t)p * 2 * max_bins + max_bins;

  // ---- gaussian smoothing, np.convolve(density, g, 'same'), left-to-right sum (:147-156) ----
  const int hw = (g.win - 1) >> 1;
  const double* gw = gauss_table + gauss_off[hw];
  double mx = -DBL_MAX;
#pragma unroll
  for (int q = 0; q < COL_BPT; ++q) {
    const int i = tid + q * COL_THREADS;
    if (i < nbins) {
      const int j0 = max(0, i - hw), j1 = min(nbins - 1, i + hw);
      double s = 0.0;
      for (int j = j0; j <= j1; ++j) s = s + dens[j] * gw[i + hw - j];
      sm_s[i] = s;
      sm_out[i] = s;
      mx = fmax(mx, s);
    }
  }
  mx = block_max_d(mx, red);   // also orders the sm_s[] writes before the reads below
  const double* sm = sm_s;
  const double hmin = mx * 0.2;                                   // :159
  const double pmin = mx * 0.05;                                  // :168
  const int dist = max(1, (int)(med / (1.5 * (double)res)));      // :163

  // ---- local maxima (plateau midpoint) + height, ascending order ----
  int npk_run = 0;
  for (int i0 = 0; i0 < nbins; i0 += COL_THREADS) {
    const int i = i0 + tid;
    int is_pk = 0, pos = 0;
    if (i >= 1 && i < nbins - 1 && sm[i - 1] < sm[i]) {
      int a = i + 1;
      while (a < nbins - 1 && sm[a] == sm[i]) ++a;
      if (sm[a] < sm[i]) {
        pos = (i + a - 1) >> 1;
        is_pk = hmin <= sm[pos] ? 1 : 0;
      }
    }
    int total;
    const int ex = pg_block_exscan(is_pk, scan_smem, &total);
    if (is_pk) {
      const int slot = npk_run + ex;
      if (slot < COL_MAX_PEAKS) { pk_pos[slot] = pos; pk_h[slot] = sm[pos]; pk_keep[slot] = 1; }
    }
    npk_run += total;
  }
  __syncthreads();
  if (npk_run > COL_MAX_PEAKS) {
    if (tid == 0) n_cols[p] = -1;
    return;
  }
  const int npk = npk_run;

  // ---- distance selection (scipy _select_by_peak_distance) ----
  // scipy visits the peaks from the highest down (equal heights: the later one first) and a visited peak that is
  // still kept removes every peak closer than `dist`.  That greedy rule has one fixed point — a peak is kept iff
  // no KEPT peak of higher priority lies within `dist` — reached here without the serial walk: one thread per
  // peak, rounds in which an undecided peak is removed when a higher-priority neighbour is kept and kept when
  // none of them is undecided any more (pk_keep: 1 undecided, 3 kept, 0 removed).  A few rounds in practice.
  for (int round = 0; round < COL_MAX_PEAKS + 1; ++round) {
    int state = tid < npk ? pk_keep[tid] : 0;
    if (state == 1) {
      const int pj = pk_pos[tid];
      const double hj = pk_h[tid];
      bool kept_near = false, open_near = false;
      for (int k = tid - 1; k >= 0 && pj - pk_pos[k] < dist; --k) {
        if (pk_h[k] > hj) {  // an earlier peak outranks this one only when strictly higher
          const int sk = pk_keep[k];
          kept_near |= sk == 3;
          open_near |= sk == 1;
        }
      }
      for (int k = tid + 1; k < npk && pk_pos[k] - pj < dist; ++k) {
        if (pk_h[k] >= hj) {  // a later peak of equal height is visited first
          const int sk = pk_keep[k];
          kept_near |= sk == 3;
          open_near |= sk == 1;
        }
      }
      state = kept_near ? 0 : (open_near ? 1 : 3);
    }
    __syncthreads();  // every thread has read the old states
    if (tid < npk) pk_keep[tid] = (unsigned char)state;
    if (!__syncthreads_or(state == 1)) break;
  }

  // ---- prominence (wlen=None): one warp per kept peak, 32 samples per step on each side ----
  __shared__ unsigned char pk_fin[COL_MAX_PEAKS];
  for (int q = warp; q < npk; q += COL_THREADS / 32) {
    if (!pk_keep[q]) { if (lane == 0) pk_fin[q] = 0; continue; }
    const int pkp = pk_pos[q];
    const double hp = sm[pkp];
    double lmin = hp, rmin = hp;
    for (int i0 = pkp; i0 >= 0; i0 -= 32) {  // leftwards while sm[i] <= hp
      const int i = i0 - lane;
      const double v = i >= 0 ? sm[i] : 0.0;
      const unsigned stop = __ballot_sync(0xffffffffu, i < 0 || v > hp);
      const int first = stop ? __ffs(stop) - 1 : 32;
      if (lane < first) lmin = fmin(lmin, v);
      if (stop) break;
    }
    for (int i0 = pkp; i0 < nbins; i0 += 32) {  // rightwards
      const int i = i0 + lane;
      const double v = i < nbins ? sm[i] : 0.0;
      const unsigned stop = __ballot_sync(0xffffffffu, i >= nbins || v > hp);
      const int first = stop ? __ffs(stop) - 1 : 32;
      if (lane < first) rmin = fmin(rmin, v);
      if (stop) break;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lmin = fmin(lmin, __shfl_xor_sync(0xffffffffu, lmin, o));
      rmin = fmin(rmin, __shfl_xor_sync(0xffffffffu, rmin, o));
    }
    if (lane == 0) pk_fin[q] = (pmin <= hp - fmax(lmin, rmin)) ? 1 : 0;
  }
  __syncthreads();
  // final ordered compaction
  int fin = 0, fpos = 0;
  if (tid < npk && pk_fin[tid]) { fin = 1; fpos = pk_pos[tid]; }
  int nfin;
  const int fex = pg_block_exscan(fin, scan_smem, &nfin);
  __syncthreads();
  if (fin) pk_pos[fex] = fpos;   // safe: every thread has read its own pk_pos[tid] above
  __syncthreads();

  // ---- centres + valley-walk widths (:176-222) ----
  if (tid < nfin && tid < max_cols) {
    const int pk = pk_pos[tid];
    int left = pk;
    if (tid > 0) {
      const int prev = pk_pos[tid - 1];
      for (int j = pk - 1; j > prev; --j) {
        if (sm[j] < sm[left]) left = j;
        if (sm[j] < hmin * 0.1) break;
      }
      if (left == pk) left = (pk + prev) >> 1;
    }
    int right = pk;
    if (tid < nfin - 1) {
      const int nxt = pk_pos[tid + 1];
      for (int j = pk + 1; j < nxt; ++j) {
        if (sm[j] < sm[right]) right = j;
        if (sm[j] < hmin * 0.1) break;
      }
      if (right == pk) right = (pk + nxt) >> 1;
    }
    double wv = (double)((right - left) * res);
    if (wv < 0.5 * med) wv = med;
    else if (wv > 2.5 * med) wv = 2.0 * med;
    centers[(int64_t)p * max_cols + tid] = pk * res;
    widths[(int64_t)p * max_cols + tid] = wv;
    if (col_hist) {
      int hb = (int)(((int64_t)pk * res * 1000) / W);
      hb = hb < 0 ? 0 : (hb > PG_COL_HIST_BINS - 1 ? PG_COL_HIST_BINS - 1 : hb);
      atomicAdd(&col_hist[hb], 1u);
    }
  }
  if (tid == 0) n_cols[p] = nfin;
}

extern "C" int pg_column_peaks(const double* boxes, const uint8_t* flags, const double* scores,
                               const int32_t* sel_idx, const int64_t* page_off, const int32_t* n_sel,
                               int32_t n_pages, const int32_t* page_wh, const double* median,
                               const double* gauss_table, const int64_t* gauss_off, int32_t max_window,
                               double min_confidence, int32_t max_cols, int32_t* centers, double* widths,
                               int32_t* n_cols, double* ws, int32_t max_bins, void* ws_spans,
                               int32_t* ws_span_counts, uint32_t* col_hist, void* stream) {
  PG_REQUIRE(n_pages >= 0, "n_pages");
  if (n_pages == 0) return PG_OK;
  PG_REQUIRE(boxes && flags && scores && page_off && page_wh && median && gauss_table && gauss_off && centers &&
                 widths && n_cols && ws && ws_spans && ws_span_counts,
             "null device pointer");
  PG_REQUIRE(max_cols > 0 && max_cols <= COL_MAX_PEAKS && max_bins > 0 && max_window > 0, "limits");
  PG_REQUIRE(n_pages <= 65535, "n_pages per launch must be <= 65535");
  PG_REQUIRE(((uintptr_t)ws_spans & 15) == 0, "ws_spans must be 16-byte aligned");
  static_assert(sizeof(DenEntry) == PG_COL_SPAN_BYTES, "PG_COL_SPAN_BYTES must match DenEntry");
  cudaStream_t s = (cudaStream_t)stream;
  DenEntry* spans = reinterpret_cast<DenEntry*>(ws_spans);
  column_prep_kernel<<<n_pages, DEN_THREADS, 0, s>>>(boxes, flags, scores, sel_idx, page_off, n_sel, page_wh, median,
                                                    max_window, min_confidence, max_bins, spans, ws_span_counts);
  PG_LAUNCH_CHECK();
  const int gx = (min(max_bins, COL_THREADS * COL_BPT) + DEN_THREADS - 1) / DEN_THREADS;
  column_density_kernel<<<dim3((unsigned)gx, (unsigned)n_pages), DEN_THREADS, 0, s>>>(
      page_off, page_wh, median, max_window, spans, ws_span_counts, ws, max_bins);
  PG_LAUNCH_CHECK();
  column_peaks_kernel<<<n_pages, COL_THREADS, 0, s>>>(page_wh, median, gauss_table, gauss_off, max_window, max_cols,
                                                     centers, widths, n_cols, ws, max_bins, col_hist);
  PG_LAUNCH_CHECK();
  return PG_OK;
}

// =============================================================================================
// K5b — per-box column assignment (derived; the reference stops at centres/widths, SURVEY §8a note)
//   column(box) = index of the column centre nearest to the box's x-centre (x0+x1)/2, first minimum
//   on ties; -1 for boxes that were not selected or pages without columns.  One thread per box.
// =============================================================================================
__global__ void __launch_bounds__(256) assign_columns_kernel(const double* __restrict__ boxes,
                                                             const int32_t* __restrict__ sel_idx,
                                                             const int64_t* __restrict__ page_off,
                                                             const int32_t* __restrict__ n_sel,
                                                             const int32_t* __restrict__ centers,
                                                             const int32_t* __restrict__ n_cols, int max_cols,
                                                             int32_t* __restrict__ col_of_box) {
  const int p = blockIdx.y;
  const int64_t base = page_off[p];
  const int m = n_sel ? n_sel[p] : (int)(page_off[p + 1] - base);
  int nc = n_cols[p];
  nc = nc < 0 ? 0 : (nc > max_cols ? max_cols : nc);
  const int32_t* cc = centers + (int64_t)p * max_cols;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < m; k += gridDim.x * blockDim.x) {
    const int64_t gi = sel_idx ? (int64_t)sel_idx[base + k] : base + k;
    const double cx = (boxes[4 * gi] + boxes[4 * gi + 2]) / 2.0;
    int best = -1;
    double bd = 0.0;
    for (int c = 0; c < nc; ++c) {
      const double dd = fabs(cx - (double)cc[c]);
      if (best < 0 || dd < bd) { best = c; bd = dd; }
    }
    col_of_box[gi] = best;
  }
}

extern "C" int pg_assign_columns(const double* boxes, const int32_t* sel_idx, const int64_t* page_off,
                                 const int32_t* n_sel, int32_t n_pages, int64_t n_boxes, const int32_t* centers,
                                 const int32_t* n_cols, int32_t max_cols, int32_t* col_of_box, void* stream) {
  PG_REQUIRE(n_pages >= 0 && n_boxes >= 0 && max_cols > 0, "sizes");
  if (n_pages == 0 || n_boxes == 0) return PG_OK;
  PG_REQUIRE(boxes && page_off && centers && n_cols && col_of_box, "null device pointer");
  PG_REQUIRE(n_pages <= 65535, "n_pages per launch must be <= 65535");
  cudaStream_t s = (cudaStream_t)stream;
  PG_CUDA_TRY(cudaMemsetAsync(col_of_box, 0xFF, (size_t)n_boxes * sizeof(int32_t), s));  // -1 everywhere
  assign_columns_kernel<<<dim3(8, (unsigned)n_pages), 256, 0, s>>>(boxes, sel_idx, page_off, n_sel, centers, n_cols,
                                                                  max_cols, col_of_box);
  PG_LAUNCH_CHECK();
  return PG_OK;
}
