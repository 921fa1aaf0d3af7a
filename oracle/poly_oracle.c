/* poly_oracle.c -- TEST INFRASTRUCTURE: CPU restatement of the DOTA result-merging NMS (fp64).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may call into this file; the product
 * (s2anet_b200/) never does.
 *
 * Follows, function by function, DOTA_devkit/polyiou/csrc/polyiou.cpp (sig :9-11, Point == :15-17, cross :19-21,
 * area :22-29, lineCross :30-39, polygon_cut :58-71, intersectArea :74-90 and :92-105, iou_poly :110-126) and
 * py_cpu_nms_poly_fast (DOTA_devkit/ResultMerge_multi_process.py:62-123).  Pinned against the reference's own
 * polyiou.cpp compiled in place (oracle/_ref/libref_polyiou.so, tests/test_poly_nms.py) and against the golden
 * vectors generated from it (tests/golden/make_golden_poly.py).
 *
 * Two places where the reference is not deterministic are fixed here and documented in DESIGN.md:
 *  - polygon_cut reads an uninitialised point when lineCross reports "no crossing" for an edge whose end points
 *    have different signs (needs |s2 - s1| <= 1e-8): the slot keeps what the previous cut of the same triangle
 *    pair left there (zero at first);
 *  - numpy's argsort()[::-1] leaves the order of equal scores unspecified: equal scores keep input order. */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define S2A_API __attribute__((visibility("default")))

typedef struct { double x, y; } ppt_t;

static int sig(double d) { return (d > 1e-8) - (d < -1e-8); }
static int same(ppt_t a, ppt_t b) { return sig(a.x - b.x) == 0 && sig(a.y - b.y) == 0; }
static double cross(ppt_t o, ppt_t a, ppt_t b) { return (a.x - o.x) * (b.y - o.y) - (b.x - o.x) * (a.y - o.y); }

static double area(ppt_t *ps, int n)
{
    ps[n] = ps[0];
    double res = 0;
    for (int i = 0; i < n; i++) res += ps[i].x * ps[i + 1].y - ps[i].y * ps[i + 1].x;
    return res / 2.0;
}

static int line_cross(ppt_t a, ppt_t b, ppt_t c, ppt_t d, ppt_t *p)
{
    double s1 = cross(a, b, c), s2 = cross(a, b, d);
    if (sig(s1) == 0 && sig(s2) == 0) return 2;
    if (sig(s2 - s1) == 0) return 0;
    p->x = (c.x * s2 - d.x * s1) / (s2 - s1);
    p->y = (c.y * s2 - d.y * s1) / (s2 - s1);
    return 1;
}

static void polygon_cut(ppt_t *p, int *n, ppt_t a, ppt_t b, ppt_t *pp)
{
    int m = 0;
    p[*n] = p[0];
    for (int i = 0; i < *n; i++) {
        if (sig(cross(a, b, p[i])) > 0) pp[m++] = p[i];
        if (sig(cross(a, b, p[i])) != sig(cross(a, b, p[i + 1]))) line_cross(a, b, p[i], p[i + 1], &pp[m++]);
    }
    *n = 0;
    for (int i = 0; i < m; i++)
        if (!i || !same(pp[i], pp[i - 1])) p[(*n)++] = pp[i];
    while (*n > 1 && same(p[*n - 1], p[0])) (*n)--;
}

static double tri_intersect(ppt_t a, ppt_t b, ppt_t c, ppt_t d)
{
    ppt_t o = { 0, 0 };
    int s1 = sig(cross(o, a, b)), s2 = sig(cross(o, c, d));
    if (s1 == 0 || s2 == 0) return 0.0;
    if (s1 == -1) { ppt_t t = a; a = b; b = t; }
    if (s2 == -1) { ppt_t t = c; c = d; d = t; }
    ppt_t p[10], pp[20];
    memset(pp, 0, sizeof(pp));
    p[0] = o; p[1] = a; p[2] = b;
    int n = 3;
    polygon_cut(p, &n, o, c, pp);
    polygon_cut(p, &n, c, d, pp);
    polygon_cut(p, &n, d, o, pp);
    double res = fabs(area(p, n));
    if (s1 * s2 == -1) res = -res;
    return res;
}

S2A_API double s2a_oracle_poly_iou(const double *p, const double *q)
{
    ppt_t ps1[5], ps2[5];
    for (int i = 0; i < 4; i++) {
        ps1[i].x = p[2 * i]; ps1[i].y = p[2 * i + 1];
        ps2[i].x = q[2 * i]; ps2[i].y = q[2 * i + 1];
    }
    if (area(ps1, 4) < 0) { ppt_t t = ps1[0]; ps1[0] = ps1[3]; ps1[3] = t; t = ps1[1]; ps1[1] = ps1[2]; ps1[2] = t; }
    if (area(ps2, 4) < 0) { ppt_t t = ps2[0]; ps2[0] = ps2[3]; ps2[3] = t; t = ps2[1]; ps2[1] = ps2[2]; ps2[2] = t; }
    ps1[4] = ps1[0];
    ps2[4] = ps2[0];
    double inter = 0;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) inter += tri_intersect(ps1[i], ps1[i + 1], ps2[j], ps2[j + 1]);
    double uni = fabs(area(ps1, 4)) + fabs(area(ps2, 4)) - inter;
    return inter / uni;
}

S2A_API void s2a_oracle_poly_iou_pairs(const double *p, const double *q, long long n, double *out)
{
    for (long long i = 0; i < n; i++) out[i] = s2a_oracle_poly_iou(p + 8 * i, q + 8 * i);
}

typedef struct { double score; long long idx; } keyed_t;
static int by_score_desc_stable(const void *a, const void *b)
{
    const keyed_t *x = a, *y = b;
    if (x->score > y->score) return -1;
    if (x->score < y->score) return 1;
    return (x->idx > y->idx) - (x->idx < y->idx);
}

/* dets [n][stride >= 9]: x0 y0 x1 y1 x2 y2 x3 y3 score.  Returns the number of kept rows; keep[] lists them in
 * the reference's order (descending score). */
S2A_API long long s2a_oracle_poly_nms(const double *dets, long long stride, long long n, double thresh, long long *keep)
{
    if (n <= 0) return 0;
    keyed_t *order = malloc(sizeof(keyed_t) * (size_t)n);
    double *hb = malloc(sizeof(double) * 5 * (size_t)n);
    char *gone = calloc((size_t)n, 1);
    for (long long i = 0; i < n; i++) {
        const double *d = dets + i * stride;
        order[i].score = d[8]; order[i].idx = i;
        double x1 = fmin(fmin(d[0], d[2]), fmin(d[4], d[6])), y1 = fmin(fmin(d[1], d[3]), fmin(d[5], d[7]));
        double x2 = fmax(fmax(d[0], d[2]), fmax(d[4], d[6])), y2 = fmax(fmax(d[1], d[3]), fmax(d[5], d[7]));
        hb[5 * i] = x1; hb[5 * i + 1] = y1; hb[5 * i + 2] = x2; hb[5 * i + 3] = y2;
        hb[5 * i + 4] = (x2 - x1 + 1) * (y2 - y1 + 1);
    }
    qsort(order, (size_t)n, sizeof(keyed_t), by_score_desc_stable);
    long long kept = 0;
    for (long long a = 0; a < n; a++) {
        if (gone[a]) continue;
        const long long i = order[a].idx;
        keep[kept++] = i;
        for (long long b = a + 1; b < n; b++) {
            if (gone[b]) continue;
            const long long j = order[b].idx;
            double xx1 = fmax(hb[5 * i], hb[5 * j]), yy1 = fmax(hb[5 * i + 1], hb[5 * j + 1]);
            double xx2 = fmin(hb[5 * i + 2], hb[5 * j + 2]), yy2 = fmin(hb[5 * i + 3], hb[5 * j + 3]);
            double w = fmax(0.0, xx2 - xx1), h = fmax(0.0, yy2 - yy1);
            double inter = w * h;
            double ovr = inter / (hb[5 * i + 4] + hb[5 * j + 4] - inter);
            if (ovr > 0) ovr = s2a_oracle_poly_iou(dets + i * stride, dets + j * stride);
            if (!(ovr <= thresh)) gone[b] = 1;
        }
    }
    free(order); free(hb); free(gone);
    return kept;
}
