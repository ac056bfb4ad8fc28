/* Plain-C restatement of the greedy NMS the reference calls.
 *
 * TEST INFRASTRUCTURE (see oracle/__init__.py) -- never linked into the product.
 *
 * Follows the published algorithm of torchvision.ops.nms (third-party dependency of the
 * reference: requirements.txt:2, call sites yolov8/tools/test.py:202 and
 * yolov8/tools/train.py:93; installed version 0.26.0; source not vendored):
 *   stable descending score sort, greedy sweep, suppress iff (double)IoU > thr,
 *   IoU = inter / (area_i + area_j - inter) in fp32, no FMA contraction
 *   (build with -ffp-contract=off), NaN never suppresses.
 * oracle_class_nms adds the reference's per-class loop (tools/test.py:181-218):
 *   strict fp32 score > conf, classes ascending, score-descending inside a class.
 */
#include <stdlib.h>
#include <string.h>

typedef struct { float score; long idx; } item_t;

static int cmp_desc(const void *a, const void *b) {
    const item_t *x = (const item_t *)a, *y = (const item_t *)b;
    if (x->score > y->score) return -1;
    if (x->score < y->score) return 1;
    return (x->idx > y->idx) - (x->idx < y->idx); /* stable: lower index first */
}

static float fmax_std(float a, float b) { return (a < b) ? b : a; }
static float fmin_std(float a, float b) { return (b < a) ? b : a; }

/* keep indices refer to positions in sel[] (or 0..n-1 when sel == NULL) */
static long greedy(const float *boxes, const float *scores, const long *sel, long n,
                   double thr, long *keep) {
    if (n <= 0) return 0;
    item_t *ord = (item_t *)malloc(sizeof(item_t) * n);
    float *area = (float *)malloc(sizeof(float) * n);
    char *sup = (char *)calloc(n, 1);
    for (long i = 0; i < n; i++) {
        long g = sel ? sel[i] : i;
        const float *b = boxes + 4 * g;
        volatile float w = b[2] - b[0];
        volatile float h = b[3] - b[1];
        area[i] = w * h;
        ord[i].score = scores[g];
        ord[i].idx = i;
    }
    qsort(ord, n, sizeof(item_t), cmp_desc);
    long k = 0;
    for (long oi = 0; oi < n; oi++) {
        long i = ord[oi].idx;
        if (sup[i]) continue;
        keep[k++] = sel ? sel[i] : i;
        const float *bi = boxes + 4 * (sel ? sel[i] : i);
        float ia = area[i];
        for (long oj = oi + 1; oj < n; oj++) {
            long j = ord[oj].idx;
            if (sup[j]) continue;
            const float *bj = boxes + 4 * (sel ? sel[j] : j);
            float xx1 = fmax_std(bi[0], bj[0]);
            float yy1 = fmax_std(bi[1], bj[1]);
            float xx2 = fmin_std(bi[2], bj[2]);
            float yy2 = fmin_std(bi[3], bj[3]);
            float w = fmax_std(0.0f, xx2 - xx1);
            float h = fmax_std(0.0f, yy2 - yy1);
            volatile float inter = w * h;
            volatile float sum = ia + area[j];
            volatile float uni = sum - inter;
            float ovr = inter / uni;
            if ((double)ovr > thr) sup[j] = 1;
        }
    }
    free(ord); free(area); free(sup);
    return k;
}

long oracle_greedy_nms(const float *boxes, const float *scores, long n, double thr, long *keep) {
    return greedy(boxes, scores, NULL, n, thr, keep);
}

long oracle_class_nms(const float *boxes, const float *scores, const int *labels, long n,
                      float conf, double thr, long *keep) {
    long k = 0;
    int maxc = -1;
    for (long i = 0; i < n; i++) if (scores[i] > conf && labels[i] > maxc) maxc = labels[i];
    long *sel = (long *)malloc(sizeof(long) * (n > 0 ? n : 1));
    for (int c = 0; c <= maxc; c++) {
        long m = 0;
        for (long i = 0; i < n; i++) if (scores[i] > conf && labels[i] == c) sel[m++] = i;
        if (m == 0) continue;
        k += greedy(boxes, scores, sel, m, thr, keep + k);
    }
    free(sel);
    return k;
}
