/* Fast CPU oracle for the class-aware greedy NMS — TEST INFRASTRUCTURE ONLY.
 *
 * Restates apply_non_max_suppression (3_combine_grids.py:80-138) and
 * calculate_iou (3_combine_grids.py:46-78) in C so that parity can be checked at
 * 10^4..10^5 boxes, where the reference's pure-Python loop needs minutes..hours.
 * "pick the first max-score box, drop same-class boxes with IoU > thr" is
 * evaluated as: visit boxes by (score desc, input position asc); a box not yet
 * suppressed is picked and suppresses every later same-class box with IoU > thr.
 * Build: see oracle/Makefile (-ffp-contract=off: no FMA contraction, doubles only).
 */
#include <stdint.h>
#include <stdlib.h>

static const double* g_scores;

static int by_score_desc(const void* pa, const void* pb) {
    int32_t a = *(const int32_t*)pa, b = *(const int32_t*)pb;
    double sa = g_scores[a], sb = g_scores[b];
    if (sa > sb) return -1;
    if (sa < sb) return 1;
    return (a > b) - (a < b);
}

static double iou(const double* a, const double* b) {
    double xl = a[0] > b[0] ? a[0] : b[0];
    double yt = a[1] > b[1] ? a[1] : b[1];
    double xr = a[2] < b[2] ? a[2] : b[2];
    double yb = a[3] < b[3] ? a[3] : b[3];
    if (xr < xl || yb < yt) return 0.0;
    double inter = (xr - xl) * (yb - yt);
    double aa = (a[2] - a[0]) * (a[3] - a[1]);
    double ab = (b[2] - b[0]) * (b[3] - b[1]);
    double uni = aa + ab - inter;
    return uni > 0 ? inter / uni : 0.0;
}

/* boxes [n,4], scores [n], classes [n]; kept_idx out [n]; returns number kept. */
int oracle_nms(const double* boxes, const double* scores, const double* classes,
               int32_t n, double thr, int32_t* kept_idx) {
    int32_t* order = (int32_t*)malloc(sizeof(int32_t) * (size_t)(n > 0 ? n : 1));
    uint8_t* dead = (uint8_t*)calloc((size_t)(n > 0 ? n : 1), 1);
    for (int32_t i = 0; i < n; ++i) order[i] = i;
    g_scores = scores;
    qsort(order, (size_t)n, sizeof(int32_t), by_score_desc);
    int32_t nk = 0;
    for (int32_t p = 0; p < n; ++p) {
        int32_t i = order[p];
        if (dead[i]) continue;
        kept_idx[nk++] = i;
        const double* bi = boxes + 4 * (size_t)i;
        double ci = classes[i];
        for (int32_t q = p + 1; q < n; ++q) {
            int32_t j = order[q];
            if (dead[j] || classes[j] != ci) continue;
            if (iou(bi, boxes + 4 * (size_t)j) > thr) dead[j] = 1;
        }
    }
    free(order);
    free(dead);
    return nk;
}
