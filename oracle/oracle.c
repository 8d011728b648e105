/*
 * oracle.c -- CPU restatement (TEST INFRASTRUCTURE, not product) of the integer / byte
 * stages of the droplet-quantification hot path of malani86/unet-DC-segmentation.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library.  The product path (unet_dc_segmentation_b200) never calls it.
 *
 * Parity pin: the reference has no tests for this path.  The restatement is
 * pinned against outputs of the reference's own functions run in the build container
 * (tests/golden/make_golden.py imports /root/reference unmodified under sys.modules shims
 * and commits the vectors), and against OpenCV / scipy where those are importable.
 *
 * What each function follows:
 *   orc_ellipse_rows      cv2.getStructuringElement(MORPH_ELLIPSE, (radius, radius))
 *                         as called at reference utils/data_loader.py:17
 *   orc_rolling_ball_u8   reference utils/data_loader.py:11-24 (per channel: MORPH_OPEN,
 *                         cv2.subtract, cv2.normalize NORM_MINMAX 0..255)
 *   orc_label4            skimage.measure.label(mask, connectivity=1) as called at
 *                         reference quantify_droplets_batch.py:82 and :86
 *   orc_quantify          reference quantify_droplets_batch.py:81-95 (label, min_area
 *                         filter, relabel, regionprops_table, micron columns)
 *   orc_resize_linear_u8  the two cv2.resize calls, reference quantify_droplets_batch.py:44 and :57
 *   orc_overlay_stencil   cv2.findContours(RETR_EXTERNAL, CHAIN_APPROX_SIMPLE) + cv2.drawContours(thickness 2),
 *                         reference quantify_droplets_batch.py:76-77, as the set of painted pixels
 *
 * Third-party arithmetic restated here (not vendored in /root/reference; requirements.txt
 * pins nothing): opencv-python (probe: 4.13.0 in this image) and scikit-image (absent in
 * this image; published semantics: 4-connectivity, labels 1..n in raster order of each
 * component's first pixel, area = pixel count, centroid = mean of integer coordinates,
 * equivalent_diameter = sqrt(4*area/pi)).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ structuring element */

/* Row i of the k x k ellipse covers columns [j1[i], j2[i]) (empty when j1 == j2).
 * OpenCV: r = k/2, c = k/2, dx = cvRound(c * sqrt((r*r - dy*dy) / (r*r))). */
int orc_ellipse_rows(int k, int *j1, int *j2)
{
    if (k < 1) return -1;
    int r = k / 2, c = k / 2;
    double inv_r2 = r ? 1.0 / ((double)r * r) : 0.0;
    for (int i = 0; i < k; ++i) {
        int dy = i - r;
        j1[i] = j2[i] = 0;
        if (abs(dy) <= r) {
            int dx = (int)nearbyint(c * sqrt(((double)r * r - (double)dy * dy) * inv_r2));
            int a = c - dx, b = c + dx + 1;
            j1[i] = a < 0 ? 0 : a;
            j2[i] = b > k ? k : b;
        }
    }
    return 0;
}

/* ------------------------------------------------------------------ morphology */

/* One flat-SE pass.  is_max = 0: erode (min), 1: dilate (max).  Both use the SAME offsets
 * (OpenCV does not reflect the element for dilation), anchor = (k/2, k/2); taps falling
 * outside the image are ignored (cv2 morphologyDefaultBorderValue). */
static void morph_pass(const uint8_t *src, uint8_t *dst, int H, int W, int k,
                       const int *j1, const int *j2, int is_max)
{
    int an = k / 2;
    for (int y = 0; y < H; ++y) {
        for (int x = 0; x < W; ++x) {
            int acc = is_max ? 0 : 255;
            for (int i = 0; i < k; ++i) {
                int yy = y + i - an;
                if (yy < 0 || yy >= H || j1[i] >= j2[i]) continue;
                int xa = x + j1[i] - an, xb = x + j2[i] - an; /* [xa, xb) */
                if (xa < 0) xa = 0;
                if (xb > W) xb = W;
                const uint8_t *row = src + (size_t)yy * W;
                if (is_max) { for (int xx = xa; xx < xb; ++xx) if (row[xx] > acc) acc = row[xx]; }
                else        { for (int xx = xa; xx < xb; ++xx) if (row[xx] < acc) acc = row[xx]; }
            }
            dst[(size_t)y * W + x] = (uint8_t)acc;
        }
    }
}

/* cv2.normalize(src, None, 0, 255, NORM_MINMAX) on a u8 plane.  OpenCV computes, in double,
 * scale = 255 * (1/(max-min)) (0 when max == min) and shift = 0 - min*scale, then convertTo
 * evaluates fma(src, float(scale), float(shift)) in fp32 and rounds half-to-even with
 * saturation.  Checked exhaustively (all 32,640 (min,max) pairs x all values) against
 * cv2 4.13.0: the fused form matches everywhere, the unfused form differs on 6,498 values. */
static void minmax_stretch(uint8_t *p, size_t n)
{
    int mn = 255, mx = 0;
    for (size_t i = 0; i < n; ++i) { if (p[i] < mn) mn = p[i]; if (p[i] > mx) mx = p[i]; }
    double scale = (mx - mn) > 0 ? 255.0 * (1.0 / (double)(mx - mn)) : 0.0;
    double shift = 0.0 - (double)mn * scale;
    float a = (float)scale, b = (float)shift;
    uint8_t lut[256];
    for (int v = 0; v < 256; ++v) {
        float f = fmaf((float)v, a, b);
        long q = lrintf(f);
        lut[v] = (uint8_t)(q < 0 ? 0 : q > 255 ? 255 : q);
    }
    for (size_t i = 0; i < n; ++i) p[i] = lut[p[i]];
}

/* image: u8 [H, W, C] interleaved (HWC); out: same shape.  Returns 0 on success. */
int orc_rolling_ball_u8(const uint8_t *image, uint8_t *out, int H, int W, int C, int radius)
{
    if (H <= 0 || W <= 0 || C <= 0 || radius < 1) return -1;
    size_t n = (size_t)H * W;
    int *j1 = (int *)malloc(sizeof(int) * radius), *j2 = (int *)malloc(sizeof(int) * radius);
    uint8_t *ch = (uint8_t *)malloc(n), *er = (uint8_t *)malloc(n), *bg = (uint8_t *)malloc(n);
    if (!j1 || !j2 || !ch || !er || !bg) { free(j1); free(j2); free(ch); free(er); free(bg); return -2; }
    orc_ellipse_rows(radius, j1, j2);
    for (int c = 0; c < C; ++c) {
        for (size_t i = 0; i < n; ++i) ch[i] = image[i * C + c];
        morph_pass(ch, er, H, W, radius, j1, j2, 0);
        morph_pass(er, bg, H, W, radius, j1, j2, 1);
        for (size_t i = 0; i < n; ++i) { int d = (int)ch[i] - (int)bg[i]; ch[i] = (uint8_t)(d < 0 ? 0 : d); }
        minmax_stretch(ch, n);
        for (size_t i = 0; i < n; ++i) out[i * C + c] = ch[i];
    }
    free(j1); free(j2); free(ch); free(er); free(bg);
    return 0;
}

/* ------------------------------------------------------------------ labelling */

static int32_t uf_find(int32_t *p, int32_t x)
{
    while (p[x] != x) { p[x] = p[p[x]]; x = p[x]; }
    return x;
}

/* 4-connected labelling of equal-valued non-zero pixels (so it serves both the first call on
 * a 0/1 mask and the second call on a label image, where it is a pure compaction).
 * labels: int32 [H, W]; returns the number of components (labels are 1..n in raster order of
 * first pixel), or <0 on error. */
int orc_label4(const int32_t *img, int32_t *labels, int H, int W)
{
    size_t n = (size_t)H * W;
    int32_t *parent = (int32_t *)malloc(sizeof(int32_t) * (n ? n : 1));
    if (!parent) return -2;
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            size_t i = (size_t)y * W + x;
            parent[i] = (int32_t)i;
            if (!img[i]) continue;
            if (x > 0 && img[i - 1] == img[i]) parent[i] = uf_find(parent, (int32_t)(i - 1));
            if (y > 0 && img[i - W] == img[i]) {
                int32_t a = uf_find(parent, (int32_t)i), b = uf_find(parent, (int32_t)(i - W));
                if (a < b) parent[b] = a; else if (b < a) parent[a] = b;
            }
        }
    /* roots are minimal raster indices, so numbering roots in raster order is skimage's order */
    int32_t next = 0;
    for (size_t i = 0; i < n; ++i) {
        if (!img[i]) { labels[i] = 0; continue; }
        int32_t r = uf_find(parent, (int32_t)i);
        if ((size_t)r == i) labels[i] = ++next;     /* first pixel of its component */
        else labels[i] = labels[r];                 /* r < i, already numbered */
    }
    free(parent);
    return next;
}

/* Reference quantify(): returns the number of droplets kept (<= capacity is required, else
 * -3) and fills per-droplet rows.  labels_out (optional) receives the final label image.
 *   area[i], sum_row[i], sum_col[i] : exact integers
 *   centroid0/1 = sum/area (f64), eq_diam = sqrt(4*area/pi) (f64)
 *   area_um2 = area / px^2, diam_um = eq_diam / px when px_per_um > 0 (else untouched). */
int orc_quantify(const uint8_t *mask, int H, int W, int64_t min_area, double px_per_um,
                 int32_t *labels_out, int capacity,
                 int64_t *area, double *centroid0, double *centroid1, double *eq_diam,
                 double *area_um2, double *diam_um)
{
    size_t n = (size_t)H * W;
    int32_t *img = (int32_t *)calloc(n ? n : 1, sizeof(int32_t));
    int32_t *lbl = (int32_t *)malloc(sizeof(int32_t) * (n ? n : 1));
    if (!img || !lbl) { free(img); free(lbl); return -2; }
    for (size_t i = 0; i < n; ++i) img[i] = mask[i] ? 1 : 0;
    int n1 = orc_label4(img, lbl, H, W);                       /* qdb:82 */
    if (n1 < 0) { free(img); free(lbl); return n1; }
    int64_t *cnt = (int64_t *)calloc((size_t)n1 + 1, sizeof(int64_t));
    for (size_t i = 0; i < n; ++i) cnt[lbl[i]]++;
    for (size_t i = 0; i < n; ++i)                             /* qdb:83-85 */
        if (lbl[i] && cnt[lbl[i]] < min_area) lbl[i] = 0;
    free(cnt);
    int n2 = orc_label4(lbl, img, H, W);                       /* qdb:86 (compaction) */
    if (n2 < 0 || n2 > capacity) { free(img); free(lbl); return n2 < 0 ? n2 : -3; }
    int64_t *sr = (int64_t *)calloc((size_t)n2 + 1, sizeof(int64_t));
    int64_t *sc = (int64_t *)calloc((size_t)n2 + 1, sizeof(int64_t));
    int64_t *ar = (int64_t *)calloc((size_t)n2 + 1, sizeof(int64_t));
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            int32_t l = img[(size_t)y * W + x];
            if (l) { ar[l]++; sr[l] += y; sc[l] += x; }
        }
    const double pi = 3.14159265358979323846;
    for (int l = 1; l <= n2; ++l) {                            /* qdb:89-94 */
        area[l - 1] = ar[l];
        centroid0[l - 1] = (double)sr[l] / (double)ar[l];
        centroid1[l - 1] = (double)sc[l] / (double)ar[l];
        eq_diam[l - 1] = sqrt(4.0 * (double)ar[l] / pi);
        if (px_per_um > 0) {
            area_um2[l - 1] = (double)ar[l] / (px_per_um * px_per_um);
            diam_um[l - 1] = eq_diam[l - 1] / px_per_um;
        }
    }
    if (labels_out) memcpy(labels_out, img, sizeof(int32_t) * n);
    free(sr); free(sc); free(ar); free(img); free(lbl);
    return n2;
}

/* ------------------------------------------------------------------ bilinear resize (u8)
 * cv2.resize(src, (dw, dh)) with INTER_LINEAR on 8-bit data, which is what BOTH resize calls of the reference
 * do: quantify_droplets_batch.py:44 passes cv2.INTER_AREA and :57 passes cv2.INTER_NEAREST in the `dst` slot, so
 * the interpolation stays at its default (pinned by tests/golden/resize.npz, generated with the
 * reference's exact call form).  OpenCV (opencv-python 4.13, imgproc resize.cpp, not vendored in /root/reference):
 *   fx = (float)((dx + 0.5) * scale - 0.5); sx = floor(fx); fx -= sx          (scale = 1 / (dsize / ssize), double)
 *   x: sx < 0 -> sx = 0, fx = 0;  sx >= sw - 1 -> sx = sw - 1, fx = 0;  taps sx and min(sx + 1, sw - 1)
 *   y: weights are NOT clamped; the two row indices sy, sy + 1 are clipped to [0, sh - 1]
 *   coefficients: short(lrintf((1 - f) * 2048)), short(lrintf(f * 2048))
 *   horizontal: S = s[sx] * a0 + s[sx + 1] * a1            (int)
 *   vertical:   dst = (((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2
 * src: u8 [sh, sw, cn] interleaved; dst: u8 [dh, dw, cn]. */
static void orc_lin_coeff(int d, int dn, int sn, int clamp_frac, int *i0, int *i1, int *a0, int *a1)
{
    double scale = 1.0 / ((double)dn / (double)sn);
    float f = (float)(((double)d + 0.5) * scale - 0.5);
    int s = (int)floorf(f);
    f -= (float)s;
    if (clamp_frac) {
        if (s < 0) { s = 0; f = 0.f; }
        if (s >= sn - 1) { s = sn - 1; f = 0.f; }
        *i0 = s; *i1 = s + 1 < sn ? s + 1 : sn - 1;
    } else {
        *i0 = s < 0 ? 0 : (s > sn - 1 ? sn - 1 : s);
        *i1 = s + 1 < 0 ? 0 : (s + 1 > sn - 1 ? sn - 1 : s + 1);
    }
    *a0 = (int)lrintf((1.f - f) * 2048.f);
    *a1 = (int)lrintf(f * 2048.f);
}

int orc_resize_linear_u8(const uint8_t *src, int sh, int sw, int cn, uint8_t *dst, int dh, int dw)
{
    if (sh <= 0 || sw <= 0 || dh <= 0 || dw <= 0 || cn <= 0) return -1;
    for (int y = 0; y < dh; ++y) {
        int y0, y1, b0, b1;
        orc_lin_coeff(y, dh, sh, 0, &y0, &y1, &b0, &b1);
        for (int x = 0; x < dw; ++x) {
            int x0, x1, a0, a1;
            orc_lin_coeff(x, dw, sw, 1, &x0, &x1, &a0, &a1);
            for (int c = 0; c < cn; ++c) {
                int S0 = src[((size_t)y0 * sw + x0) * cn + c] * a0 + src[((size_t)y0 * sw + x1) * cn + c] * a1;
                int S1 = src[((size_t)y1 * sw + x0) * cn + c] * a0 + src[((size_t)y1 * sw + x1) * cn + c] * a1;
                int v = (((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2;
                dst[((size_t)y * dw + x) * cn + c] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
            }
        }
    }
    return 0;
}


/* ---------------------------------------------------------------------------------------------
 * Overlay stencil: which pixels reference quantify_droplets_batch.py:76-77 paints,
 *     cnts, _ = cv2.findContours(mask, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
 *     cv2.drawContours(img, cnts, -1, (0, 255, 0), 2)
 * restated without border following (OpenCV 4.x semantics; tests/test_oracle_golden.py pins it against cv2):
 *  - RETR_EXTERNAL keeps the outer borders of the 8-connected non-zero components that face the background
 *    connected (4-connectivity, through zero pixels) to the image frame; a component inside a hole of another
 *    has a hole border as parent and is dropped.
 *  - The points of such a border are the component's pixels with a 4-neighbour in that outer background
 *    (Suzuki-Abe border points for 8-connected foreground); CHAIN_APPROX_SIMPLE only drops collinear ones.
 *  - drawContours thickness 2: ThickLine per polyline segment = a filled radius-1 circle (a plus) at both end
 *    points + FillConvexPoly of the band p +- dp, |dp| = 1.  For horizontal / vertical segments that is the pixel
 *    row / column on either side; for diagonal segments dp = (+-0.707, -+0.707) and the polygon OUTLINE, drawn
 *    with a fixed-point line, rounds onto the two diagonals one step to either side of the segment.
 *    A diagonal step p -> q exists where p, q are diagonal neighbours and the 4-neighbour they share on one side
 *    is outer background.
 * out[y*W+x] = 1 where painted.  Returns 0, or -1 on bad arguments / allocation failure. */
int orc_overlay_stencil(const uint8_t *mask, int H, int W, uint8_t *out)
{
    if (H <= 0 || W <= 0 || !mask || !out) return -1;
    const int P = 3, PH = H + 2 * P, PW = W + 2 * P;        /* padded frame: zeros = outer background */
    uint8_t *fg = (uint8_t *)calloc((size_t)PH * PW, 1);
    uint8_t *outer = (uint8_t *)calloc((size_t)PH * PW, 1);
    uint8_t *ct = (uint8_t *)calloc((size_t)PH * PW, 1);
    int *stack = (int *)malloc(sizeof(int) * (size_t)PH * PW);
    if (!fg || !outer || !ct || !stack) { free(fg); free(outer); free(ct); free(stack); return -1; }
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) fg[(y + P) * PW + x + P] = mask[y * W + x] != 0;
    /* flood the zero pixels from the frame corner, 4-connectivity */
    int sp = 0;
    stack[sp++] = 0;
    outer[0] = 1;
    while (sp) {
        const int i = stack[--sp], y = i / PW, x = i % PW;
        const int ny[4] = {y - 1, y + 1, y, y}, nx[4] = {x, x, x - 1, x + 1};
        for (int k = 0; k < 4; ++k) {
            if (ny[k] < 0 || ny[k] >= PH || nx[k] < 0 || nx[k] >= PW) continue;
            const int j = ny[k] * PW + nx[k];
            if (!fg[j] && !outer[j]) { outer[j] = 1; stack[sp++] = j; }
        }
    }
    for (int y = 1; y < PH - 1; ++y)
        for (int x = 1; x < PW - 1; ++x) {
            const int i = y * PW + x;
            ct[i] = fg[i] && (outer[i - PW] || outer[i + PW] || outer[i - 1] || outer[i + 1]);
        }
#define ORC_PAINT(yy, xx)                                                                             \
    do {                                                                                              \
        const int y_ = (yy) - P, x_ = (xx) - P;                                                       \
        if (y_ >= 0 && y_ < H && x_ >= 0 && x_ < W) out[y_ * W + x_] = 1;                             \
    } while (0)
    for (int i = 0; i < H * W; ++i) out[i] = 0;
    for (int y = 1; y < PH - 2; ++y)
        for (int x = 1; x < PW - 1; ++x) {
            const int i = y * PW + x;
            if (ct[i]) { ORC_PAINT(y, x); ORC_PAINT(y - 1, x); ORC_PAINT(y + 1, x); ORC_PAINT(y, x - 1); ORC_PAINT(y, x + 1); }
            if (!fg[i]) continue;
            if (fg[i + PW + 1] && (outer[i + 1] || outer[i + PW])) {        /* step to (y+1, x+1) */
                ORC_PAINT(y - 1, x + 1); ORC_PAINT(y + 1, x - 1); ORC_PAINT(y, x + 2); ORC_PAINT(y + 2, x);
            }
            if (fg[i + PW - 1] && (outer[i - 1] || outer[i + PW])) {        /* step to (y+1, x-1) */
                ORC_PAINT(y + 1, x + 1); ORC_PAINT(y - 1, x - 1); ORC_PAINT(y + 2, x); ORC_PAINT(y, x - 2);
            }
        }
#undef ORC_PAINT
    free(fg); free(outer); free(ct); free(stack);
    return 0;
}
