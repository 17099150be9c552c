// pg_boxes.cu — K2 edge filter, K3 cross-tile NMS merge, K4 width median, K5 column centres.
//
// All four are fp64, bit-exact restatements of the reference's Python-double arithmetic
// (pg_math.h), batched over pages through CSR offsets so that one launch serves a whole
// shard.  None of them is HBM-bound (a page's boxes are a few hundred KB); they are organised
// so that the dependent chain per page is short and pages run side by side on different SMs.
#include <cfloat>

#include "pg_common.cuh"

// =============================================================================================
// K2 — edge-touch filter (+ fused cell->page translation)
//   reference: translate_coordinates_to_original 1_doclayout_bboxes.py:484-511,
//              is_box_touching_internal_edge 2_edge_box_filter.py:44-90,
//              filter_grid_info 2_edge_box_filter.py:206-217 (kept order = input order)
// One CTA per page; order-preserving compaction by block scan.
// =============================================================================================
__global__ void __launch_bounds__(256) edge_filter_kernel(const double* __restrict__ boxes, int local,
                                                          const int32_t* __restrict__ box_cell,
                                                          const double* __restrict__ cells,
                                                          const int32_t* __restrict__ page_wh,
                                                          const int64_t* __restrict__ page_off, double thr,
                                                          double* __restrict__ boxes_out, uint8_t* __restrict__ keep,
                                                          int32_t* __restrict__ kept_idx, int32_t* __restrict__ n_kept) {
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
  edge_filter_kernel<<<n_pages, 256, 0, (cudaStream_t)stream>>>(boxes, boxes_are_local, box_cell, cells, page_wh,
                                                                page_off, threshold, boxes_page_out, keep, kept_idx,
                                                                n_kept);
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
//   A  bin      per page, counting-sort the boxes into a 64x64 grid of centre cells (x-major)
//               so that runs of 32 consecutive boxes ("blocks") are spatially compact; write a
//               blocked SoA copy and each block's bounding box;
//   B  count    per block I, count blocks J of the same page whose bounding boxes intersect
//               (IoU > thr >= 0 needs a non-empty intersection, the reference's own early-out);
//   C  scan     exclusive scan of the counts -> entry offsets (workspace overflow -> status);
//   D  fill     per candidate pair (I,J): lane i of the warp holds box i of I, the 32 boxes of J
//               are broadcast from shared memory; lane i accumulates a 32-bit mask of the boxes
//               of J that would suppress it;
//   E  resolve  per page, Jacobi rounds over two bit-words per block (kept / undecided): an
//               undecided box becomes suppressed if a suppressor is kept, kept if none of its
//               suppressors is still undecided.  Each round decides at least the best undecided
//               box, typical depth is 3-6 rounds;
//   F  emit     rank the kept boxes by priority (counting rank, shared-memory tiles) and write
//               global indices in pick order.
// =============================================================================================
constexpr int NMS_GX = 64, NMS_GY = 64, NMS_CELLS = NMS_GX * NMS_GY;

struct NmsWs {
  int64_t* stats;      // [8]: status, candidate pairs, rounds, box pairs tested
  int32_t* sorted_pos; // [N]   spatial order -> local position k
  int32_t* cellid;     // [N]
  double* sx0; double* sy0; double* sx1; double* sy1; double* sarea; double* sscore; double* scls;  // [NB*32]
  int32_t* skpos;      // [NB*32] local position, -1 = padding lane
  double* bbox;        // [NB*4]
  int32_t* blk_page;   // [NB]
  int32_t* cand_cnt;   // [NB]
  int64_t* cand_off;   // [NB+1]
  uint32_t* st_kept;   // [2*NB]
  uint32_t* st_undec;  // [2*NB]
  double* kscore;      // [N]
  int32_t* kpos;       // [N]
  int32_t* ent_j;      // [E]
  uint32_t* ent_mask;  // [E*32]
  int64_t nb_cap, ent_cap;
};

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
  w.sx0 = (double*)take((size_t)nb * 32 * 8);
  w.sy0 = (double*)take((size_t)nb * 32 * 8);
  w.sx1 = (double*)take((size_t)nb * 32 * 8);
  w.sy1 = (double*)take((size_t)nb * 32 * 8);
  w.sarea = (double*)take((size_t)nb * 32 * 8);
  w.sscore = (double*)take((size_t)nb * 32 * 8);
  w.scls = (double*)take((size_t)nb * 32 * 8);
  w.skpos = (int32_t*)take((size_t)nb * 32 * 4);
  w.bbox = (double*)take((size_t)nb * 4 * 8);
  w.blk_page = (int32_t*)take((size_t)nb * 4);
  w.cand_cnt = (int32_t*)take((size_t)nb * 4);
  w.cand_off = (int64_t*)take((size_t)(nb + 1) * 8);
  w.st_kept = (uint32_t*)take((size_t)nb * 2 * 4);
  w.st_undec = (uint32_t*)take((size_t)nb * 2 * 4);
  w.kscore = (double*)take((size_t)n * 8);
  w.kpos = (int32_t*)take((size_t)n * 4);
  w.ent_j = (int32_t*)take((size_t)ecap * 4);
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

// ---- A: bin --------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) nms_bin_kernel(const double* __restrict__ boxes,
                                                       const double* __restrict__ scores,
                                                       const double* __restrict__ classes,
                                                       const int32_t* __restrict__ sel_idx,
                                                       const int64_t* __restrict__ page_off,
                                                       const int32_t* __restrict__ n_sel, int n_pages, NmsWs ws) {
  __shared__ int counts[NMS_CELLS];
  __shared__ double red[4][32];
  __shared__ int scan_smem[34];
  const int p = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const PageSpan sp = page_span(page_off, n_sel, p);
  const int64_t blk_next = (p + 1 < n_pages) ? (page_off[p + 1] >> 5) + p + 1 : ws.nb_cap;
  for (int64_t b = sp.blk0 + sp.nb + tid; b < blk_next; b += blockDim.x) ws.blk_page[b] = -1;

  // extent of the box centres
  double mnx = DBL_MAX, mny = DBL_MAX, mxx = -DBL_MAX, mxy = -DBL_MAX;
  for (int k = tid; k < sp.m; k += blockDim.x) {
    const int64_t gi = sel_idx ? (int64_t)sel_idx[sp.base + k] : sp.base + k;
    const double cx = (boxes[4 * gi] + boxes[4 * gi + 2]) * 0.5, cy = (boxes[4 * gi + 1] + boxes[4 * gi + 3]) * 0.5;
    mnx = fmin(mnx, cx); mxx = fmax(mxx, cx); mny = fmin(mny, cy); mxy = fmax(mxy, cy);
  }
  mnx = warp_min_d(mnx); mny = warp_min_d(mny); mxx = warp_max_d(mxx); mxy = warp_max_d(mxy);
  if (lane == 0) { red[0][warp] = mnx; red[1][warp] = mny; red[2][warp] = mxx; red[3][warp] = mxy; }
  for (int c = tid; c < NMS_CELLS; c += blockDim.x) counts[c] = 0;
  __syncthreads();
  if (warp == 0) {
    double a = warp_min_d(red[0][lane]), b = warp_min_d(red[1][lane]);
    double c = warp_max_d(red[2][lane]), d = warp_max_d(red[3][lane]);
    if (lane == 0) { red[0][0] = a; red[1][0] = b; red[2][0] = c; red[3][0] = d; }
  }
  __syncthreads();
  mnx = red[0][0]; mny = red[1][0]; mxx = red[2][0]; mxy = red[3][0];
  const double sx = (mxx > mnx) ? (double)NMS_GX / (mxx - mnx) : 0.0;
  const double sy = (mxy > mny) ? (double)NMS_GY / (mxy - mny) : 0.0;

  // cell ids + histogram
  for (int k = tid; k < sp.m; k += blockDim.x) {
    const int64_t gi = sel_idx ? (int64_t)sel_idx[sp.base + k] : sp.base + k;
    const double cx = (boxes[4 * gi] + boxes[4 * gi + 2]) * 0.5, cy = (boxes[4 * gi + 1] + boxes[4 * gi + 3]) * 0.5;
    const double fx = (cx - mnx) * sx, fy = (cy - mny) * sy;
    int gx = (fx >= 0.0) ? (fx < (double)(NMS_GX - 1) ? (int)fx : NMS_GX - 1) : 0;
    int gy = (fy >= 0.0) ? (fy < (double)(NMS_GY - 1) ? (int)fy : NMS_GY - 1) : 0;
    const int cell = gx * NMS_GY + gy;  // x-major: consecutive cells walk down a column
    ws.cellid[sp.base + k] = cell;
    atomicAdd(&counts[cell], 1);
  }
  __syncthreads();
  {  // exclusive scan of the 4096 cell counts (4 per thread)
    const int c0 = counts[4 * tid], c1 = counts[4 * tid + 1], c2 = counts[4 * tid + 2], c3 = counts[4 * tid + 3];
    int total;
    const int ex = pg_block_exscan(c0 + c1 + c2 + c3, scan_smem, &total);
    counts[4 * tid] = ex; counts[4 * tid + 1] = ex + c0; counts[4 * tid + 2] = ex + c0 + c1;
    counts[4 * tid + 3] = ex + c0 + c1 + c2;
  }
  __syncthreads();
  // stable scatter by one warp (deterministic spatial order)
  if (warp == 0) {
    for (int c = 0; c < sp.m; c += 32) {
      const int k = c + lane;
      const bool act = k < sp.m;
      const int cell = act ? ws.cellid[sp.base + k] : -1 - lane;
      const unsigned grp = __match_any_sync(0xffffffffu, cell);
      const int leader = __ffs(grp) - 1;
      const int rank = __popc(grp & ((1u << lane) - 1u));
      int basepos = 0;
      if (act && lane == leader) { basepos = counts[cell]; counts[cell] = basepos + __popc(grp); }
      basepos = __shfl_sync(0xffffffffu, basepos, leader);
      if (act) ws.sorted_pos[sp.base + basepos + rank] = k;
      __syncwarp();
    }
  }
  __syncthreads();
  // blocked SoA copy + block bounding boxes
  for (int b = warp; b < sp.nb; b += (blockDim.x >> 5)) {
    const int pos = b * 32 + lane;
    const bool valid = pos < sp.m;
    const int64_t slot = (sp.blk0 + b) * 32 + lane;
    double x0 = 0, y0 = 0, x1 = 0, y1 = 0, sc = 0, cl = 0;
    int k = -1;
    if (valid) {
      k = ws.sorted_pos[sp.base + pos];
      const int64_t gi = sel_idx ? (int64_t)sel_idx[sp.base + k] : sp.base + k;
      x0 = boxes[4 * gi]; y0 = boxes[4 * gi + 1]; x1 = boxes[4 * gi + 2]; y1 = boxes[4 * gi + 3];
      sc = scores[gi]; cl = classes[gi];
    }
    ws.sx0[slot] = x0; ws.sy0[slot] = y0; ws.sx1[slot] = x1; ws.sy1[slot] = y1;
    ws.sarea[slot] = pg_box_area(x0, y0, x1, y1);
    ws.sscore[slot] = sc; ws.scls[slot] = cl; ws.skpos[slot] = k;
    const double bx0 = warp_min_d(valid ? fmin(x0, x1) : DBL_MAX), by0 = warp_min_d(valid ? fmin(y0, y1) : DBL_MAX);
    const double bx1 = warp_max_d(valid ? fmax(x0, x1) : -DBL_MAX), by1 = warp_max_d(valid ? fmax(y0, y1) : -DBL_MAX);
    if (lane == 0) {
      double* bb = ws.bbox + 4 * (sp.blk0 + b);
      bb[0] = bx0; bb[1] = by0; bb[2] = bx1; bb[3] = by1;
      ws.blk_page[sp.blk0 + b] = p;
    }
  }
}

__device__ __forceinline__ bool bbox_hit(const double* a, const double* b) {
  return !(b[2] < a[0] || a[2] < b[0] || b[3] < a[1] || a[3] < b[1]);
}

// ---- B: count / D: fill (same traversal) ---------------------------------------------------
template <bool FILL>
__global__ void __launch_bounds__(256) nms_pairs_kernel(const int64_t* __restrict__ page_off,
                                                        const int32_t* __restrict__ n_sel, NmsWs ws, double thr,
                                                        int all_pairs) {
  __shared__ double jbox[8][7][32];
  __shared__ int jk[8][32];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t I = (int64_t)blockIdx.x * (blockDim.x >> 5) + wib;
  if (I >= ws.nb_cap) return;
  if (FILL && ws.stats[0] != PG_OK) return;
  const int p = ws.blk_page[I];
  if (p < 0) {
    if (!FILL && lane == 0) ws.cand_cnt[I] = 0;
    return;
  }
  const PageSpan sp = page_span(page_off, n_sel, p);
  double bbI[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) bbI[q] = ws.bbox[4 * I + q];

  // lane i <-> box i of block I
  double ix0 = 0, iy0 = 0, ix1 = 0, iy1 = 0, iarea = 0, iscore = 0, icls = 0;
  int ik = -1;
  int64_t e = 0;
  if (FILL) {
    const int64_t slot = I * 32 + lane;
    ix0 = ws.sx0[slot]; iy0 = ws.sy0[slot]; ix1 = ws.sx1[slot]; iy1 = ws.sy1[slot];
    iarea = ws.sarea[slot]; iscore = ws.sscore[slot]; icls = ws.scls[slot]; ik = ws.skpos[slot];
    e = ws.cand_off[I];
  }
  int cnt = 0;
  for (int j0 = 0; j0 < sp.nb; j0 += 32) {
    const int64_t J = sp.blk0 + j0 + lane;
    bool hit = false;
    if (j0 + lane < sp.nb) hit = all_pairs || bbox_hit(bbI, ws.bbox + 4 * J);
    unsigned hits = __ballot_sync(0xffffffffu, hit);
    cnt += __popc(hits);
    if (FILL) {
      while (hits) {
        const int b = __ffs(hits) - 1;
        hits &= hits - 1;
        const int64_t Jb = sp.blk0 + j0 + b;
        const int64_t js = Jb * 32 + lane;
        __syncwarp();
        jbox[wib][0][lane] = ws.sx0[js]; jbox[wib][1][lane] = ws.sy0[js]; jbox[wib][2][lane] = ws.sx1[js];
        jbox[wib][3][lane] = ws.sy1[js]; jbox[wib][4][lane] = ws.sarea[js]; jbox[wib][5][lane] = ws.sscore[js];
        jbox[wib][6][lane] = ws.scls[js]; jk[wib][lane] = ws.skpos[js];
        __syncwarp();
        uint32_t mask = 0;
        if (ik >= 0) {
#pragma unroll 4
          for (int jj = 0; jj < 32; ++jj) {
            const int kj = jk[wib][jj];
            const double sj = jbox[wib][5][jj];
            // j must outrank i: higher score, or equal score and earlier pooled position (:112)
            const bool outranks = (sj > iscore) || (sj == iscore && kj < ik);
            if (kj >= 0 && outranks && jbox[wib][6][jj] == icls) {
              const double v = pg_iou(jbox[wib][0][jj], jbox[wib][1][jj], jbox[wib][2][jj], jbox[wib][3][jj],
                                      jbox[wib][4][jj], ix0, iy0, ix1, iy1, iarea);
              if (v > thr) mask |= 1u << jj;
            }
          }
        }
        const bool any = __any_sync(0xffffffffu, mask != 0);
        ws.ent_mask[e * 32 + lane] = mask;
        if (lane == 0) ws.ent_j[e] = any ? (int32_t)(Jb - sp.blk0) : -1;
        ++e;
      }
    }
  }
  if (!FILL) {
    if (lane == 0) ws.cand_cnt[I] = cnt;
  } else if (lane == 0 && cnt) {
    atomicAdd((unsigned long long*)&ws.stats[3], (unsigned long long)cnt * 1024ull);
  }
}

// ---- C: scan -------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) nms_scan_kernel(NmsWs ws) {
  __shared__ int scan_smem[34];
  __shared__ long long carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int64_t b0 = 0; b0 < ws.nb_cap; b0 += blockDim.x) {
    const int64_t b = b0 + threadIdx.x;
    const int v = b < ws.nb_cap ? ws.cand_cnt[b] : 0;
    int total;
    const int ex = pg_block_exscan(v, scan_smem, &total);
    const long long carry = carry_s;
    if (b < ws.nb_cap) ws.cand_off[b] = carry + ex;
    __syncthreads();
    if (threadIdx.x == 0) carry_s = carry + total;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    ws.cand_off[ws.nb_cap] = carry_s;
    ws.stats[1] = carry_s;
    ws.stats[0] = (carry_s > ws.ent_cap) ? PG_ERR_WORKSPACE : PG_OK;
  }
}

// ---- E: resolve ----------------------------------------------------------------------------
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
  int cur = 0, rounds = 0;
  while (true) {
    const int64_t rc = (int64_t)cur * nbc, rn = (int64_t)(cur ^ 1) * nbc;
    int pending = 0;
    for (int b = warp; b < sp.nb; b += nwarps) {
      const int64_t I = sp.blk0 + b;
      const uint32_t u = undec[rc + I], k = kept[rc + I];
      if (u == 0u) {
        if (lane == 0) { undec[rn + I] = 0u; kept[rn + I] = k; }
        continue;
      }
      const bool mine = (u >> lane) & 1u;
      bool sup = false, wait = false;
      const int64_t e0 = ws.cand_off[I], e1 = ws.cand_off[I + 1];
      for (int64_t e = e0; e < e1; ++e) {
        const int jb = ws.ent_j[e];
        if (jb < 0) continue;
        const uint32_t m = mine ? ws.ent_mask[e * 32 + lane] : 0u;
        const uint32_t kj = kept[rc + sp.blk0 + jb], uj = undec[rc + sp.blk0 + jb];
        if (m & kj) sup = true;
        else if (m & uj) wait = true;
      }
      const unsigned bk = __ballot_sync(0xffffffffu, mine && !sup && !wait);
      const unsigned bs = __ballot_sync(0xffffffffu, mine && sup);
      const uint32_t un = u & ~(bk | bs);
      if (lane == 0) { kept[rn + I] = k | bk; undec[rn + I] = un; }
      pending |= (un != 0u);
    }
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
    int ex = pg_block_exscan(__popc(k), scan_smem, &total);
    uint32_t bits = k;
    int64_t dst = sp.base + running + ex;
    while (bits) {
      const int l = __ffs(bits) - 1;
      bits &= bits - 1;
      const int64_t slot = (sp.blk0 + b) * 32 + l;
      ws.kscore[dst] = ws.sscore[slot];
      ws.kpos[dst] = ws.skpos[slot];
      ++dst;
    }
    running += total;
  }
  if (tid == 0) {
    n_kept[p] = running;
    atomicMax((unsigned long long*)&ws.stats[2], (unsigned long long)rounds);
  }
}

// ---- F: emit -------------------------------------------------------------------------------
constexpr int EMIT_TILE = 1024;
__global__ void __launch_bounds__(EMIT_TILE) nms_emit_kernel(const int32_t* __restrict__ sel_idx,
                                                             const int64_t* __restrict__ page_off, NmsWs ws,
                                                             const int32_t* __restrict__ n_kept,
                                                             int32_t* __restrict__ kept_idx) {
  __shared__ double ts[EMIT_TILE];
  __shared__ int tk[EMIT_TILE];
  const int p = blockIdx.x, tid = threadIdx.x;
  const int K = n_kept[p];
  if (K <= 0) return;
  const int64_t base = page_off[p];
  for (int i0 = blockIdx.y * EMIT_TILE; i0 < K; i0 += gridDim.y * EMIT_TILE) {
    const int i = i0 + tid;
    const bool act = i < K;
    const double si = act ? ws.kscore[base + i] : 0.0;
    const int ki = act ? ws.kpos[base + i] : 0;
    int rank = 0;
    for (int t0 = 0; t0 < K; t0 += EMIT_TILE) {
      __syncthreads();
      if (t0 + tid < K) { ts[tid] = ws.kscore[base + t0 + tid]; tk[tid] = ws.kpos[base + t0 + tid]; }
      __syncthreads();
      const int n = min(EMIT_TILE, K - t0);
      if (act) {
#pragma unroll 8
        for (int j = 0; j < n; ++j) {
          const double sj = ts[j];
          rank += (sj > si || (sj == si && tk[j] < ki)) ? 1 : 0;
        }
      }
    }
    if (act) kept_idx[base + rank] = sel_idx ? sel_idx[base + ki] : (int32_t)(base + ki);
  }
}

extern "C" int pg_nms_merge(const double* boxes, const double* scores, const double* classes,
                            const int32_t* sel_idx, const int64_t* page_off, const int32_t* n_sel,
                            int32_t n_pages, int64_t n_boxes, int32_t max_boxes_per_page, double iou_threshold,
                            int32_t* kept_idx, int32_t* n_kept, void* workspace, size_t workspace_bytes,
                            void* stream) {
  PG_REQUIRE(n_pages >= 0 && n_boxes >= 0, "sizes");
  if (n_pages == 0) return PG_OK;
  PG_REQUIRE(boxes && scores && classes && page_off && kept_idx && n_kept && workspace, "null device pointer");
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
  cudaStream_t s = (cudaStream_t)stream;
  PG_CUDA_TRY(cudaMemsetAsync(ws.stats, 0, 8 * sizeof(int64_t), s));
  const int all_pairs = !(iou_threshold >= 0.0);  // thr < 0: disjoint boxes (IoU 0) suppress too
  nms_bin_kernel<<<n_pages, 1024, 0, s>>>(boxes, scores, classes, sel_idx, page_off, n_sel, n_pages, ws);
  PG_LAUNCH_CHECK();
  const unsigned pair_grid = (unsigned)((ws.nb_cap + 7) / 8);
  nms_pairs_kernel<false><<<pair_grid, 256, 0, s>>>(page_off, n_sel, ws, iou_threshold, all_pairs);
  PG_LAUNCH_CHECK();
  nms_scan_kernel<<<1, 1024, 0, s>>>(ws);
  PG_LAUNCH_CHECK();
  nms_pairs_kernel<true><<<pair_grid, 256, 0, s>>>(page_off, n_sel, ws, iou_threshold, all_pairs);
  PG_LAUNCH_CHECK();
  nms_resolve_kernel<<<n_pages, 1024, 0, s>>>(page_off, n_sel, ws, n_kept);
  PG_LAUNCH_CHECK();
  const int64_t mb = max_boxes_per_page > 0 ? max_boxes_per_page : n_boxes;
  int gy = (int)((mb + EMIT_TILE - 1) / EMIT_TILE);
  gy = gy < 1 ? 1 : (gy > 64 ? 64 : gy);
  nms_emit_kernel<<<dim3((unsigned)n_pages, (unsigned)gy), EMIT_TILE, 0, s>>>(sel_idx, page_off, ws, n_kept, kept_idx);
  PG_LAUNCH_CHECK();
  return PG_OK;
}

extern "C" int pg_nms_stats(const void* workspace, int64_t stats[4]) {
  PG_REQUIRE(workspace && stats, "null pointer");
  int64_t h[8];
  PG_CUDA_TRY(cudaMemcpy(h, workspace, sizeof(h), cudaMemcpyDeviceToHost));
  for (int i = 0; i < 4; ++i) stats[i] = h[i];
  return PG_OK;
}

// =============================================================================================
// K4 — plain_text width bins + median
//   reference: bin_widths 4_extract_median_widths.py:49-80 (sequential leader binning, joins the
//   smallest-key bin within the margin), calculate_median_width :82-101 (np.median), width
//   extraction :135-141.
// One warp per page.  32 widths are matched against the sorted bin keys at once (binary search
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

__global__ void __launch_bounds__(32) width_median_kernel(const double* __restrict__ boxes,
                                                          const uint8_t* __restrict__ flags,
                                                          const int32_t* __restrict__ sel_idx,
                                                          const int64_t* __restrict__ page_off,
                                                          const int32_t* __restrict__ n_sel,
                                                          const int32_t* __restrict__ page_wh, double margin_pct,
                                                          double* __restrict__ median, int32_t* __restrict__ n_bins,
                                                          double* ws_keys, int32_t* ws_counts, uint32_t* width_hist) {
  const int p = blockIdx.x, lane = threadIdx.x;
  const int64_t base = page_off[p];
  const int m = n_sel ? n_sel[p] : (int)(page_off[p + 1] - base);
  double* keys = ws_keys + base;
  int* counts = ws_counts + base;
  const double margin = (double)page_wh[2 * p] * (margin_pct / 100.0);  // :64
  int nb = 0;
  for (int c = 0; c < m; c += 32) {
    const int k = c + lane;
    bool act = false;
    double w = 0.0;
    if (k < m) {
      const int64_t gi = sel_idx ? (int64_t)sel_idx[base + k] : base + k;
      if (flags[gi] & PG_FLAG_PLAIN_TEXT) {
        act = true;
        w = boxes[4 * gi + 2] - boxes[4 * gi];  // :139
        if (width_hist) {
          const int hb = w >= 0.0 ? (w < (double)(PG_WIDTH_HIST_BINS - 1) ? (int)w : PG_WIDTH_HIST_BINS - 1) : 0;
          atomicAdd(&width_hist[hb], 1u);
        }
      }
    }
    unsigned pend = __ballot_sync(0xffffffffu, act);
    if (!pend) continue;
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
  // median over keys repeated by count; keys are already ascending
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
                               const int64_t* page_off, const int32_t* n_sel, int32_t n_pages,
                               const int32_t* page_wh, double min_margin_percent, double* median,
                               int32_t* n_bins, double* ws_keys, int32_t* ws_counts, uint32_t* width_hist,
                               void* stream) {
  PG_REQUIRE(n_pages >= 0, "n_pages");
  if (n_pages == 0) return PG_OK;
  PG_REQUIRE(boxes && flags && page_off && page_wh && median && n_bins && ws_keys && ws_counts, "null device pointer");
  width_median_kernel<<<n_pages, 32, 0, (cudaStream_t)stream>>>(boxes, flags, sel_idx, page_off, n_sel, page_wh,
                                                                 min_margin_percent, median, n_bins, ws_keys,
                                                                 ws_counts, width_hist);
  PG_LAUNCH_CHECK();
  return PG_OK;
}

// =============================================================================================
// K5 — column centres
//   reference: find_column_centers 5_detect_column_centers.py:91-224 (+ scipy find_peaks step
//   order: local maxima -> height -> distance -> prominence; np.convolve 'same').
// One CTA per page, one thread per density bin.  Each bin accumulates its own contributions in
// box order (the reference's `density[bin] += w` order), so the density map is bit-identical to
// numpy's; the smoothing sum runs left to right.
// =============================================================================================
constexpr int COL_THREADS = 1024;
constexpr int COL_BPT = 8;        // bins per thread -> up to 8192 bins
constexpr int COL_MAX_PEAKS = 1024;

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
    const double* __restrict__ boxes, const uint8_t* __restrict__ flags, const double* __restrict__ scores,
    const int32_t* __restrict__ sel_idx, const int64_t* __restrict__ page_off, const int32_t* __restrict__ n_sel,
    const int32_t* __restrict__ page_wh, const double* __restrict__ median, const double* __restrict__ gauss_table,
    const int64_t* __restrict__ gauss_off, int max_window, double min_conf, int max_cols,
    int32_t* __restrict__ centers, double* __restrict__ widths, int32_t* __restrict__ n_cols, double* ws_all,
    int max_bins, uint32_t* col_hist) {
  __shared__ int4 elist[COL_THREADS];
  __shared__ int scan_smem[34];
  __shared__ double red[33];
  __shared__ int pk_pos[COL_MAX_PEAKS];
  __shared__ double pk_h[COL_MAX_PEAKS];
  __shared__ unsigned char pk_keep[COL_MAX_PEAKS];
  __shared__ int s_npk;

  const int p = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t base = page_off[p];
  const int m = n_sel ? n_sel[p] : (int)(page_off[p + 1] - base);
  const int W = page_wh[2 * p], Hh = page_wh[2 * p + 1];
  const double med = median[p];
  // process_page guards (5_detect_column_centers.py:361-364, 381-383)
  if (!(med > 0.0) || W <= 0 || Hh <= 0) {
    if (tid == 0) n_cols[p] = 0;
    return;
  }
  const int res = max(1, W / 1000);                 // :120
  const int nbins = W / res + 1;                    // :121
  int win = max(5, (int)(med / (4.0 * (double)res)));  // :147
  if ((win & 1) == 0) win += 1;
  if (nbins > max_bins || nbins > COL_THREADS * COL_BPT || win > max_window || win > nbins) {
    if (tid == 0) n_cols[p] = -1;
    return;
  }
  double* dens = ws_all + (int64_t)p * 2 * max_bins;
  double* sm = dens + max_bins;
  const double lo_w = 0.33 * med, hi_w = 2.0 * med;  // :131

  // ---- density map -------------------------------------------------------------------------
  double d[COL_BPT];
#pragma unroll
  for (int q = 0; q < COL_BPT; ++q) d[q] = 0.0;
  const int wb0 = warp * 32, wstride = COL_THREADS;  // bins owned: tid + q*1024
  for (int c0 = 0; c0 < m; c0 += COL_THREADS) {
    const int k = c0 + tid;
    int acc = 0;
    int4 ent = make_int4(0, 0, 0, 0);
    if (k < m) {
      const int64_t gi = sel_idx ? (int64_t)sel_idx[base + k] : base + k;
      if ((flags[gi] & (PG_FLAG_PLAIN_TEXT | PG_FLAG_TITLE)) && scores[gi] >= min_conf) {  // :110-112
        const int x1 = (int)boxes[4 * gi], x2 = (int)boxes[4 * gi + 2];                    // :127 int() truncation
        const int bw = x2 - x1;
        if (lo_w <= (double)bw && (double)bw <= hi_w) {
          const int left = max(0, pg_floordiv(x1, res));               // :133
          const int right = min(nbins - 1, pg_floordiv(x2, res));      // :134
          const int center = pg_floordiv(x1 + x2, 2 * res);            // :137
          if (left <= right) { acc = 1; ent = make_int4(left, right, center, 0); }
        }
      }
    }
    int total;
    const int ex = pg_block_exscan(acc, scan_smem, &total);
    if (acc) elist[ex] = ent;
    __syncthreads();
    for (int e = 0; e < total; ++e) {
      const int4 en = elist[e];
#pragma unroll
      for (int q = 0; q < COL_BPT; ++q) {
        const int q0 = wb0 + q * wstride;           // first bin of this warp's q-th slab
        if (q0 < nbins && en.x <= q0 + 31 && en.y >= q0) {   // warp-uniform reject
          const int b = q0 + lane;
          if (b >= en.x && b <= en.y) d[q] = d[q] + pg_density_weight(b, en.x, en.y, en.z);  // :140-144
        }
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int q = 0; q < COL_BPT; ++q) {
    const int b = tid + q * wstride;
    if (b < nbins) dens[b] = d[q];
  }
  __syncthreads();

  // ---- gaussian smoothing, np.convolve(density, g, 'same') ----------------------------------
  const int hw = (win - 1) >> 1;
  const double* g = gauss_table + gauss_off[hw];
  double mx = -DBL_MAX;
#pragma unroll
  for (int q = 0; q < COL_BPT; ++q) {
    const int i = tid + q * wstride;
    if (i < nbins) {
      const int j0 = max(0, i - hw), j1 = min(nbins - 1, i + hw);
      double s = 0.0;
      for (int j = j0; j <= j1; ++j) s = s + dens[j] * g[i + hw - j];
      sm[i] = s;
      mx = fmax(mx, s);
    }
  }
  mx = block_max_d(mx, red);   // also orders the sm[] writes before the reads below
  const double hmin = mx * 0.2;                                   // :159
  const double pmin = mx * 0.05;                                  // :168
  const int dist = max(1, (int)(med / (1.5 * (double)res)));      // :163

  // ---- local maxima (plateau midpoint) + height, ascending order ----------------------------
  if (tid == 0) s_npk = 0;
  __syncthreads();
  int npk_run = 0;
  bool overflow = false;
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
  if (npk_run > COL_MAX_PEAKS) overflow = true;
  __syncthreads();
  if (overflow) {
    if (tid == 0) n_cols[p] = -1;
    return;
  }
  const int npk = npk_run;

  // ---- distance selection (scipy _select_by_peak_distance), warp 0 --------------------------
  if (warp == 0 && npk > 0) {
    // visited flags live in bit 1 of pk_keep
    for (int it = 0; it < npk; ++it) {
      double bh = -DBL_MAX;
      int bi = -1;
      for (int j = lane; j < npk; j += 32) {
        if (pk_keep[j] == 1) {  // kept and not yet visited
          const double h = pk_h[j];
          if (h > bh || (h == bh && j > bi)) { bh = h; bi = j; }
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double oh = __shfl_xor_sync(0xffffffffu, bh, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (oi >= 0 && (bi < 0 || oh > bh || (oh == bh && oi > bi))) { bh = oh; bi = oi; }
      }
      if (bi < 0) break;
      __syncwarp();
      if (lane == 0) {
        pk_keep[bi] = 3;  // kept + visited
        const int pj = pk_pos[bi];
        for (int k = bi - 1; k >= 0 && pj - pk_pos[k] < dist; --k) pk_keep[k] = 0;
        for (int k = bi + 1; k < npk && pk_pos[k] - pj < dist; ++k) pk_keep[k] = 0;
      }
      __syncwarp();
    }
  }
  __syncthreads();

  // ---- prominence (wlen=None) + final ordered compaction ------------------------------------
  int fin = 0, fpos = 0;
  if (tid < npk && pk_keep[tid]) {
    const int pkp = pk_pos[tid];
    const double hp = sm[pkp];
    double lmin = hp, rmin = hp;
    for (int i = pkp; i >= 0 && sm[i] <= hp; --i) lmin = fmin(lmin, sm[i]);
    for (int i = pkp; i < nbins && sm[i] <= hp; ++i) rmin = fmin(rmin, sm[i]);
    const double prom = hp - fmax(lmin, rmin);
    if (pmin <= prom) { fin = 1; fpos = pkp; }
  }
  int nfin;
  const int fex = pg_block_exscan(fin, scan_smem, &nfin);
  __syncthreads();
  if (fin) pk_pos[fex] = fpos;   // safe: every thread has read its own pk_pos[tid] above
  __syncthreads();

  // ---- centres + valley-walk widths (:176-222) ----------------------------------------------
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
                               int32_t* n_cols, double* ws, int32_t max_bins, uint32_t* col_hist, void* stream) {
  PG_REQUIRE(n_pages >= 0, "n_pages");
  if (n_pages == 0) return PG_OK;
  PG_REQUIRE(boxes && flags && scores && page_off && page_wh && median && gauss_table && gauss_off && centers &&
                 widths && n_cols && ws,
             "null device pointer");
  PG_REQUIRE(max_cols > 0 && max_cols <= COL_MAX_PEAKS && max_bins > 0 && max_window > 0, "limits");
  column_peaks_kernel<<<n_pages, COL_THREADS, 0, (cudaStream_t)stream>>>(
      boxes, flags, scores, sel_idx, page_off, n_sel, page_wh, median, gauss_table, gauss_off, max_window,
      min_confidence, max_cols, centers, widths, n_cols, ws, max_bins, col_hist);
  PG_LAUNCH_CHECK();
  return PG_OK;
}
