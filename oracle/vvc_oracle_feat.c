/*
 * TEST INFRASTRUCTURE -- CPU restatement ("oracle") of the orig-only texture measures of the reference:
 *   a16  EncCu::updateCtuDataISlice / xCalcHADs8x8_ISlice   (EL/EncCu.cpp:564-675, caller EL/EncSlice.cpp:1276-1298)
 *   a17  the fork's FAST_ALGORITHM feature block            (EL/EncCu.cpp:72-164 helpers, :816-1138 features)
 * See vvc_oracle.h for who may use it.
 *
 * Parity status.  a16: PINNED (tests/golden 'H' records of the unmodified reference, tests/test_oracle_features.py).
 * a17: PARITY UNPINNED against the reference binary -- FAST_ALGORITHM=1 cannot be built in this image (needs the
 * OpenCV C++ library and two .pkl models that are not shipped).  The restatement follows the source
 * line by line; every OpenCV primitive it stands on is cross-checked against the Python cv2 wheel (same OpenCV
 * kernels) by tests/test_oracle_features.py: convertTo(CV_8U), filter2D (correlation, BORDER_REFLECT_101, 8-bit
 * saturation), meanStdDev (population, then squared again), addWeighted (what `A/4 + B/4 + ...` lowers to).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "vvc_oracle.h"

/* ---------------------------------------------------------------------------------------------- a16 */

/* xCalcHADs8x8_ISlice (EL/EncCu.cpp:564-654): 8x8 Hadamard of the ORIGINAL samples, AC part only */
static int hads8x8_islice(const int16_t* org, int stride)
{
  int m[8][8], t[8][8], i, j, s, sum = 0;
  for (i = 0; i < 8; i++) for (j = 0; j < 8; j++) m[i][j] = org[i * stride + j];
  for (s = 4; s >= 1; s >>= 1) {                          /* rows */
    for (i = 0; i < 8; i++)
      for (j = 0; j < 8; j++) t[i][j] = (j & s) ? m[i][j - s] - m[i][j] : m[i][j] + m[i][j + s];
    memcpy(m, t, sizeof(m));
  }
  for (s = 4; s >= 1; s >>= 1) {                          /* columns */
    for (i = 0; i < 8; i++)
      for (j = 0; j < 8; j++) t[i][j] = (i & s) ? m[i - s][j] - m[i][j] : m[i][j] + m[i + s][j];
    memcpy(m, t, sizeof(m));
  }
  for (i = 0; i < 8; i++) for (j = 0; j < 8; j++) sum += abs(m[i][j]);
  sum -= abs(m[0][0]);
  return (sum + 2) >> 2;
}

/* EncCu::updateCtuDataISlice (:656-675) over the CTU grid of calCostSliceI (EL/EncSlice.cpp:1287-1293) */
void orc_ctu_hads_islice(const int16_t* orig, int stride, int pic_w, int pic_h, int ctu, int32_t* out)
{
  int cx, cy, n = 0, x, y;
  for (cy = 0; cy < pic_h; cy += ctu)
    for (cx = 0; cx < pic_w; cx += ctu) {
      const int w = pic_w - cx < ctu ? pic_w - cx : ctu, h = pic_h - cy < ctu ? pic_h - cy : ctu;
      int sum = 0;
      for (y = 0; y + 8 <= h; y += 8)
        for (x = 0; x + 8 <= w; x += 8) sum += hads8x8_islice(orig + (size_t)(cy + y) * stride + cx + x, stride);
      out[n++] = sum;
    }
}

/* ---------------------------------------------------------------------------------------------- a17 */

/* Mat::convertTo(CV_8U) of int data: saturate_cast<uchar> (EL/EncCu.cpp:152, :938) */
static int sat8(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }

static void load_u8(const int16_t* orig, int stride, int x, int y, int w, int h, uint8_t* px)
{
  int i, j;
  for (i = 0; i < h; i++) for (j = 0; j < w; j++) px[i * w + j] = (uint8_t)sat8(orig[(size_t)(y + i) * stride + x + j]);
}

/* cv::meanStdDev on one channel, then `stddev[0] * stddev[0]` (EL/EncCu.cpp:155-158, :944-950, :1056...):
 * OpenCV: mean = sum * (1/N); sd = sqrt(max(sqsum * (1/N) - mean*mean, 0)).  Sizes are powers of two, so every
 * product/difference before the square root is exact in double.                                              */
static double var_from_sums(double sum, double sqsum, int n)
{
  const double scale = 1.0 / n;
  const double mean = sum * scale;
  double v = sqsum * scale - mean * mean;
  double sd;
  if (v < 0.0) v = 0.0;
  sd = sqrt(v);
  return sd * sd;
}

static double var_u8(const uint8_t* px, int stride, int x0, int y0, int w, int h)
{
  double s = 0, q = 0;
  int i, j;
  for (i = 0; i < h; i++) for (j = 0; j < w; j++) { const int v = px[(y0 + i) * stride + x0 + j]; s += v; q += v * v; }
  return var_from_sums(s, q, w * h);
}

/* cv::filter2D(Pixel, dst, CV_8U, kern): correlation, anchor at the centre, BORDER_REFLECT_101, saturate to uchar */
static void filter3x3_u8(const uint8_t* px, int w, int h, const int k[9], uint8_t* dst)
{
  int i, j, a, b;
  for (i = 0; i < h; i++)
    for (j = 0; j < w; j++) {
      int acc = 0;
      for (a = -1; a <= 1; a++)
        for (b = -1; b <= 1; b++) {
          int yy = i + a, xx = j + b;
          if (yy < 0) yy = -yy;
          if (yy >= h) yy = 2 * h - 2 - yy;
          if (xx < 0) xx = -xx;
          if (xx >= w) xx = 2 * w - 2 - xx;
          acc += k[(a + 1) * 3 + (b + 1)] * px[yy * w + xx];
        }
      dst[i * w + j] = (uint8_t)sat8(acc);
    }
}

/* cv::addWeighted on 8-bit data: saturate_cast<uchar>(cvRound(a*alpha + b*beta)), round half to even.  The operands
 * here are multiples of 1/4 below 512, so the float arithmetic of the library is exact.                        */
static int add_weighted_u8(int a, double alpha, int b, double beta)
{
  return sat8((int)nearbyint(a * alpha + b * beta));      /* default rounding mode = to nearest even */
}

/* EncCu::get_madp (EL/EncCu.cpp:73-134): mean absolute difference to the 3 / 5 / 8 existing neighbours, integer division */
static void madp_map(const uint8_t* px, int w, int h, int* madp)
{
  int i, j, a, b;
  for (i = 0; i < h; i++)
    for (j = 0; j < w; j++) {
      int acc = 0, cnt = 0;
      for (a = -1; a <= 1; a++)
        for (b = -1; b <= 1; b++) {
          if ((a == 0 && b == 0) || i + a < 0 || i + a >= h || j + b < 0 || j + b >= w) continue;
          acc += abs((int)px[(i + a) * w + j + b] - (int)px[i * w + j]);
          cnt++;
        }
      madp[i * w + j] = acc / cnt;
    }
}

static int cmp_int(const void* a, const void* b) { return *(const int*)a - *(const int*)b; }

/* variance-of-variances of 2, 3 or 4 parts with the reference's integer truncations (EL/EncCu.cpp:1053-1095) */
static int sccd(const int* v, int n)
{
  int i, mean = 0, acc = 0;
  for (i = 0; i < n; i++) mean += v[i];
  mean /= n;
  for (i = 0; i < n; i++) acc += (v[i] - mean) * (v[i] - mean);
  return acc / n;
}

void orc_features(const int16_t* orig, int stride, const vvcb_feat_job* job, vvcb_feat_result* out)
{
  const int x = job->cu.x, y = job->cu.y, w = job->cu.w, h = job->cu.h, n = w * h;
  uint8_t* px = (uint8_t*)calloc((size_t)n, 1);
  uint8_t* g[4];
  int* madp = (int*)malloc(sizeof(int) * n);
  int* f = out->f;
  int i, k, gmax = 0;
  double gsum[4], G[4], gra, var, nmse, ms = 0, mq = 0;
  static const int kern[4][9] = {
    { -1, 0, 1, -2, 0, 2, -1, 0, 1 },        /* kern_H   :997  */
    { 1, 2, 1, 0, 0, 0, -1, -2, -1 },        /* kern_V   :1000 */
    { 0, 1, 2, -1, 0, 1, -2, -1, 0 },        /* kern_45  :1008 */
    { 2, 1, 0, 1, 0, -1, 0, -1, -2 } };      /* kern_135 :1011 */
  int ncc[5], nqt[5], nmt[5], v[4];
  memset(out, 0, sizeof(*out));
  load_u8(orig, stride, x, y, w, h, px);

  for (k = 0; k < job->n_neighbours; k++) {              /* get_context (:137-163) */
    const vvcb_feat_cu* c = &job->nb[k];
    uint8_t* q = (uint8_t*)malloc((size_t)c->w * c->h);
    load_u8(orig, stride, c->x, c->y, c->w, c->h, q);
    ncc[k] = (int)var_u8(q, c->w, 0, 0, c->w, c->h);
    nqt[k] = c->qt_depth;
    nmt[k] = c->mt_depth;
    free(q);
  }

  madp_map(px, w, h, madp);
  var = var_u8(px, w, 0, 0, w, h);                       /* :944-946 */
  for (i = 0; i < n; i++) { ms += madp[i]; mq += (double)madp[i] * madp[i]; }
  nmse = var_from_sums(ms, mq, n);                       /* :947-950 */

  for (k = 0; k < 4; k++) {
    g[k] = (uint8_t*)malloc(n);
    filter3x3_u8(px, w, h, kern[k], g[k]);               /* :1003-1004, :1016-1017 */
    gsum[k] = 0;
    for (i = 0; i < n; i++) gsum[k] += g[k][i];
    G[k] = gsum[k] / n;                                  /* :1019-1022 */
  }
  gra = (G[0] + G[1] + G[2] + G[3]) / 4;                 /* :1023 */
  /* :1027  minMaxIdx(Gra_H/4 + Gra_V/4 + Gra_45/4 + Gra_135/4): cv::MatExpr folds `A*a + B*b` of two scaled matrices
   * into one addWeighted, and every further `+ C/4` into addWeighted(previous 8-bit result, 1, C, 0.25)          */
  for (i = 0; i < n; i++) {
    int t = add_weighted_u8(g[0][i], 0.25, g[1][i], 0.25);
    t = add_weighted_u8(t, 1.0, g[2][i], 0.25);
    t = add_weighted_u8(t, 1.0, g[3][i], 0.25);
    if (t > gmax) gmax = t;
  }

  f[0] = h; f[1] = w; f[2] = job->cu.qt_depth; f[3] = job->cu.mt_depth;     /* :1098-1101 */
  f[4] = (int)G[0]; f[5] = (int)G[1]; f[6] = (int)G[2]; f[7] = (int)G[3];
  f[8] = (int)gra; f[9] = gmax;
  f[10] = (int)var; f[11] = (int)nmse;
  if (job->n_neighbours > 0) {                            /* :966-995 */
    const int m = job->n_neighbours;
    int s;
    qsort(ncc, m, sizeof(int), cmp_int); qsort(nqt, m, sizeof(int), cmp_int); qsort(nmt, m, sizeof(int), cmp_int);
    for (s = 0, k = 0; k < m; k++) s += ncc[k];
    f[12] = ncc[m - 1]; f[13] = ncc[0]; f[14] = s / m;
    for (s = 0, k = 0; k < m; k++) s += nqt[k];
    f[15] = nqt[m - 1]; f[16] = nqt[0]; f[17] = s / m;
    for (s = 0, k = 0; k < m; k++) s += nmt[k];
    f[18] = nmt[m - 1]; f[19] = nmt[0]; f[20] = s / m;
  }
  v[0] = (int)var_u8(px, w, 0, 0, w, h / 2); v[1] = (int)var_u8(px, w, 0, h / 2, w, h / 2);            /* BH :1053-1058 */
  f[21] = sccd(v, 2);
  v[0] = (int)var_u8(px, w, 0, 0, w / 2, h); v[1] = (int)var_u8(px, w, w / 2, 0, w / 2, h);            /* BV :1060-1065 */
  f[22] = sccd(v, 2);
  v[0] = (int)var_u8(px, w, 0, 0, w, h / 4); v[1] = (int)var_u8(px, w, 0, h / 4, w, h / 2);            /* TH :1067-1074 */
  v[2] = (int)var_u8(px, w, 0, 3 * h / 4, w, h / 4);
  f[23] = sccd(v, 3);
  v[0] = (int)var_u8(px, w, 0, 0, w / 4, h); v[1] = (int)var_u8(px, w, w / 4, 0, w / 2, h);            /* TV :1077-1084 */
  v[2] = (int)var_u8(px, w, 3 * w / 4, 0, w / 4, h);
  f[24] = sccd(v, 3);
  v[0] = (int)var_u8(px, w, 0, 0, w / 2, h / 2);     v[1] = (int)var_u8(px, w, w / 2, 0, w / 2, h / 2); /* QT :1086-1095 */
  v[2] = (int)var_u8(px, w, 0, h / 2, w / 2, h / 2); v[3] = (int)var_u8(px, w, w / 2, h / 2, w / 2, h / 2);
  f[25] = sccd(v, 4);
  f[26] = f[10] < f[13] ? 0 : (f[10] > f[12] ? 2 : 1);                      /* :1127-1138 */
  out->valid = job->n_neighbours >= 3;                                      /* :935 */
  for (k = 0; k < 4; k++) free(g[k]);
  free(px); free(madp);
}

void orc_features_batch(const int16_t* orig, int stride, const vvcb_feat_job* jobs, int n, vvcb_feat_result* out)
{
  int i;
  for (i = 0; i < n; i++) orc_features(orig, stride, &jobs[i], &out[i]);
}
