/* CPU restatement (TEST INFRASTRUCTURE ONLY) of the image-ingest arithmetic of the reference's load_images()
 * (fast3r/dust3r/utils/image.py:76-159): PIL.Image.resize(new_size, LANCZOS | BICUBIC) on 8-bit RGB, then center crop,
 * torchvision ToTensor + Normalize(0.5, 0.5).
 *
 * The resize itself lives in a third-party dependency that is not under /root/reference: Pillow (requirements.txt does
 * not pin it; this image ships Pillow 11/12, `python -c "import PIL; print(PIL.__version__)"`).  What follows restates
 * Pillow's published algorithm (src/libImaging/Resample.c: precompute_coeffs, normalize_coeffs_8bpc,
 * ImagingResampleHorizontal_8bpc / Vertical_8bpc) and is pinned bit-exactly against Pillow itself by
 * tests/test_ingest_cpu.py (every size / filter combination the reference can produce).
 *
 *   - per output coordinate xx: center = (xx + 0.5) * scale, support = filter_support * max(scale, 1),
 *     taps xmin = (int)(center - support + 0.5) clamped to 0, xmax = (int)(center + support + 0.5) clamped to inSize,
 *     weights filter((x + xmin - center + 0.5) / max(scale, 1)) normalised to sum 1 (double precision);
 *   - 8-bit images: weights are rounded to fixed point with 22 fractional bits (round half away from zero), the
 *     accumulator starts at 1 << 21, the result is (acc >> 22) clamped to 0..255;
 *   - two passes: horizontal first (into an 8-bit intermediate image), then vertical.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define PRECISION_BITS (32 - 8 - 2)

static double sinc_filter(double x) {
  if (x == 0.0) return 1.0;
  x = x * M_PI;
  return sin(x) / x;
}
static double lanczos_filter(double x) {
  if (-3.0 <= x && x < 3.0) return sinc_filter(x) * sinc_filter(x / 3);
  return 0.0;
}
static double bicubic_filter(double x) {
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
  if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
  return 0.0;
}

/* filter: 0 = BICUBIC (support 2), 1 = LANCZOS (support 3) */
int f3r_oracle_ksize(int in_size, int out_size, int filter) {
  double scale = (double)in_size / out_size;
  double filterscale = scale < 1.0 ? 1.0 : scale;
  double support = (filter ? 3.0 : 2.0) * filterscale;
  return (int)ceil(support) * 2 + 1;
}

/* bounds: [out_size][2] = (xmin, count); kk: [out_size][ksize] fixed-point weights */
void f3r_oracle_coeffs(int in_size, int out_size, int filter, int32_t* bounds, int32_t* kk) {
  double (*filt)(double) = filter ? lanczos_filter : bicubic_filter;
  double scale = (double)in_size / out_size;
  double filterscale = scale < 1.0 ? 1.0 : scale;
  double support = (filter ? 3.0 : 2.0) * filterscale;
  int ksize = (int)ceil(support) * 2 + 1;
  double* k = (double*)malloc(sizeof(double) * ksize);
  for (int xx = 0; xx < out_size; xx++) {
    double center = (xx + 0.5) * scale;
    double ww = 0.0, ss = 1.0 / filterscale;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    for (int x = 0; x < xmax; x++) {
      double w = filt((x + xmin - center + 0.5) * ss);
      k[x] = w;
      ww += w;
    }
    for (int x = 0; x < ksize; x++) {
      double v = 0.0;
      if (x < xmax) v = (ww != 0.0) ? k[x] / ww : k[x];
      kk[xx * ksize + x] = v < 0 ? (int32_t)(-0.5 + v * (1 << PRECISION_BITS)) : (int32_t)(0.5 + v * (1 << PRECISION_BITS));
    }
    bounds[xx * 2 + 0] = xmin;
    bounds[xx * 2 + 1] = xmax;
  }
  free(k);
}

static uint8_t clip8(int32_t v) {
  v >>= PRECISION_BITS; /* arithmetic shift, like Pillow's lookup index */
  return v < 0 ? 0 : v > 255 ? 255 : (uint8_t)v;
}

/* in: [h][w][3] uint8, out: [oh][ow][3] uint8 */
void f3r_oracle_resize_rgb8(const uint8_t* in, int h, int w, uint8_t* out, int oh, int ow, int filter) {
  const uint8_t* src = in;
  uint8_t* tmp = NULL;
  if (ow != w) { /* horizontal pass */
    int ks = f3r_oracle_ksize(w, ow, filter);
    int32_t* b = (int32_t*)malloc(sizeof(int32_t) * 2 * ow);
    int32_t* kk = (int32_t*)malloc(sizeof(int32_t) * (size_t)ks * ow);
    f3r_oracle_coeffs(w, ow, filter, b, kk);
    tmp = (uint8_t*)malloc((size_t)h * ow * 3);
    for (int y = 0; y < h; y++)
      for (int xx = 0; xx < ow; xx++) {
        int xmin = b[2 * xx], n = b[2 * xx + 1];
        const int32_t* k = kk + (size_t)xx * ks;
        for (int c = 0; c < 3; c++) {
          int32_t ss = 1 << (PRECISION_BITS - 1);
          for (int x = 0; x < n; x++) ss += (int32_t)in[((size_t)y * w + x + xmin) * 3 + c] * k[x];
          tmp[((size_t)y * ow + xx) * 3 + c] = clip8(ss);
        }
      }
    free(b); free(kk);
    src = tmp;
  }
  if (oh != h) { /* vertical pass */
    int ks = f3r_oracle_ksize(h, oh, filter);
    int32_t* b = (int32_t*)malloc(sizeof(int32_t) * 2 * oh);
    int32_t* kk = (int32_t*)malloc(sizeof(int32_t) * (size_t)ks * oh);
    f3r_oracle_coeffs(h, oh, filter, b, kk);
    for (int yy = 0; yy < oh; yy++) {
      int ymin = b[2 * yy], n = b[2 * yy + 1];
      const int32_t* k = kk + (size_t)yy * ks;
      for (int x = 0; x < ow; x++)
        for (int c = 0; c < 3; c++) {
          int32_t ss = 1 << (PRECISION_BITS - 1);
          for (int y = 0; y < n; y++) ss += (int32_t)src[((size_t)(y + ymin) * ow + x) * 3 + c] * k[y];
          out[((size_t)yy * ow + x) * 3 + c] = clip8(ss);
        }
    }
    free(b); free(kk);
  } else {
    memcpy(out, src, (size_t)oh * ow * 3);
  }
  free(tmp);
}

/* crop box (left, top, cw, ch) of a [h][w][3] uint8 image -> fp32 [3][ch][cw]: ToTensor (x / 255) then Normalize
 * ((t - 0.5) / 0.5), both in fp32 like torchvision (fast3r/dust3r/utils/image.py:32) */
void f3r_oracle_crop_normalize(const uint8_t* in, int h, int w, int left, int top, int cw, int ch, float* out) {
  (void)h;
  for (int c = 0; c < 3; c++)
    for (int y = 0; y < ch; y++)
      for (int x = 0; x < cw; x++) {
        float t = (float)in[((size_t)(y + top) * w + x + left) * 3 + c] / 255.0f;
        out[((size_t)c * ch + y) * cw + x] = (t - 0.5f) / 0.5f;
      }
}
