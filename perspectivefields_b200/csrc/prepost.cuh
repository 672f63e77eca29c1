// HBM-bound ends of the path: Pillow-exact uint8 resize + normalisation (pre) and resample-to-original (post).
#pragma once
#include <math.h>

#include <vector>

#include "common.cuh"

namespace pf {

// =====================================================================================================
// Pre-process.  Reference: perspectivefields.py:38-46 (PIL.Image.resize((320,320), BILINEAR) on uint8) and
// :234-236 ((x - pixel_mean) / pixel_std).  Pillow's resampler (src/libImaging/Resample.c, third-party) is an
// antialiased separable triangle filter with 22-bit fixed-point coefficients, horizontal pass first, each pass
// rounded to uint8.  The coefficient tables are built on the host in double precision exactly as Pillow's
// precompute_coeffs/normalize_coeffs_8bpc do; the kernel is pure integer arithmetic and therefore bit-exact.
constexpr int kPrecisionBits = 32 - 8 - 2;

struct ResampleTable {  // host-side, for one (in_size -> 320) axis
  int in_size = 0, ksize = 0;
  std::vector<int> bounds;  // [320][2] = (xmin, count)
  std::vector<int> coeffs;  // [320][ksize]
};

inline ResampleTable make_resample_table(int in_size, int out_size) {
  ResampleTable t;
  t.in_size = in_size;
  const double scale = (double)in_size / (double)out_size;
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = 1.0 * filterscale;  // bilinear (triangle) filter support = 1
  t.ksize = (int)ceil(support) * 2 + 1;
  t.bounds.assign((size_t)out_size * 2, 0);
  t.coeffs.assign((size_t)out_size * t.ksize, 0);
  std::vector<double> k((size_t)t.ksize);
  const double ss = 1.0 / filterscale;
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = (xx + 0.5) * scale;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    const int n = xmax - xmin;
    double ww = 0.0;
    for (int x = 0; x < n; ++x) {
      double a = (x + xmin - center + 0.5) * ss;
      if (a < 0.0) a = -a;
      const double w = a < 1.0 ? 1.0 - a : 0.0;
      k[x] = w;
      ww += w;
    }
    for (int x = 0; x < n; ++x) {
      double v = k[x];
      if (ww != 0.0) v /= ww;
      t.coeffs[(size_t)xx * t.ksize + x] = v < 0 ? (int)(-0.5 + v * (1 << kPrecisionBits)) : (int)(0.5 + v * (1 << kPrecisionBits));
    }
    t.bounds[2 * xx] = xmin;
    t.bounds[2 * xx + 1] = n;
  }
  return t;
}

struct PreImage {          // per image, device-visible
  long long offset;        // byte offset of the HWC uint8 image in the input blob
  int H, W;
  int ksx, ksy;            // table widths
  const int* bx; const int* cx;   // horizontal tables (depend on W)
  const int* by; const int* cy;   // vertical tables   (depend on H)
};

constexpr int kPreRows = 8;  // output rows per block

// grid = (320 / kPreRows, n_images), block = 320 threads (one per output column).
// dyn smem = rows_needed * 320 * 3 bytes for the horizontally resampled input rows of this tile.
// out: [n, 320, 320, 4] fp32 NHWC, channels (b, g, r, 0), value = (u8 - mean[c]) / std[c].
__global__ void __launch_bounds__(kNet) preprocess_kernel(const unsigned char* __restrict__ blob, const PreImage* __restrict__ imgs, float* __restrict__ out,
                                                          float m0, float m1, float m2, float s0, float s1, float s2, int max_rows) {
  extern __shared__ unsigned char s_h[];  // [rows][320][3]
  const PreImage im = imgs[blockIdx.y];
  const int oy0 = blockIdx.x * kPreRows;
  const int x = threadIdx.x;
  const unsigned char* src = blob + im.offset;
  const int xmin = im.bx[2 * x], xn = im.bx[2 * x + 1];
  const int* kx = im.cx + (long long)x * im.ksx;
  for (int r0 = 0; r0 < kPreRows;) {
    // process as many output rows as fit in smem (normally all kPreRows at once)
    int rcount = 0;
    const int in_first = im.by[2 * (oy0 + r0)];
    int in_last = in_first;
    while (r0 + rcount < kPreRows) {
      const int o = oy0 + r0 + rcount;
      const int last = im.by[2 * o] + im.by[2 * o + 1];
      if (last - in_first > max_rows && rcount > 0) break;
      in_last = last;
      ++rcount;
    }
    const int nrows = in_last - in_first;
    // horizontal pass: input rows [in_first, in_last) -> uint8 [nrows][320][3]
    for (int r = 0; r < nrows; ++r) {
      const unsigned char* row = src + ((long long)(in_first + r) * im.W + xmin) * 3;
      int a0 = 1 << (kPrecisionBits - 1), a1 = a0, a2 = a0;
      for (int t = 0; t < xn; ++t) {
        const int k = kx[t];
        a0 += row[3 * t] * k; a1 += row[3 * t + 1] * k; a2 += row[3 * t + 2] * k;
      }
      unsigned char* d = s_h + (r * kNet + x) * 3;
      d[0] = (unsigned char)min(max(a0 >> kPrecisionBits, 0), 255);
      d[1] = (unsigned char)min(max(a1 >> kPrecisionBits, 0), 255);
      d[2] = (unsigned char)min(max(a2 >> kPrecisionBits, 0), 255);
    }
    __syncthreads();
    // vertical pass
    for (int rr = 0; rr < rcount; ++rr) {
      const int o = oy0 + r0 + rr;
      const int ymin = im.by[2 * o] - in_first, yn = im.by[2 * o + 1];
      const int* ky = im.cy + (long long)o * im.ksy;
      int a0 = 1 << (kPrecisionBits - 1), a1 = a0, a2 = a0;
      for (int t = 0; t < yn; ++t) {
        const int k = ky[t];
        const unsigned char* d = s_h + ((ymin + t) * kNet + x) * 3;
        a0 += d[0] * k; a1 += d[1] * k; a2 += d[2] * k;
      }
      const float v0 = (float)min(max(a0 >> kPrecisionBits, 0), 255);
      const float v1 = (float)min(max(a1 >> kPrecisionBits, 0), 255);
      const float v2 = (float)min(max(a2 >> kPrecisionBits, 0), 255);
      reinterpret_cast<float4*>(out)[((long long)blockIdx.y * kNet + o) * kNet + x] =
          make_float4((v0 - m0) / s0, (v1 - m1) / s1, (v2 - m2) / s2, 0.f);
    }
    __syncthreads();
    r0 += rcount;
  }
}

// Lower entry (perspectivefields.py:223-236 called directly): images already resized, fp32 CHW [n,3,320,320].
__global__ void __launch_bounds__(256) normalize_chw_kernel(const float* __restrict__ in, float* __restrict__ out, int n,
                                                            float m0, float m1, float m2, float s0, float s1, float s2) {
  const long long total = (long long)n * kNet * kNet;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long b = i / (kNet * kNet), p = i % (kNet * kNet);
  const float* s = in + b * 3 * kNet * kNet + p;
  reinterpret_cast<float4*>(out)[i] = make_float4((s[0] - m0) / s0, (s[kNet * kNet] - m1) / s1, (s[2 * kNet * kNet] - m2) / s2, 0.f);
}

// =====================================================================================================
// Post-process.  Reference: gravity_head.py:237-261, latitude_head.py:195-219, utils/utils.py:483-507.
//   gravity : vec * (W/320, H/320) -> bilinear (align_corners=False, no antialias) to (H, W) -> F.normalize(dim=0)
//   latitude: bilinear to (H, W) -> asin -> rad2deg          (regression)   |  bilinear of decoded degrees (classification)
// ATen upsample_bilinear2d: scale = (float)320 / out; src = scale*(dst+0.5)-0.5, clamped at 0; i1 = i0 + (i0 < 319).
struct PostImage {
  int H, W;
  long long g_off;  // float offset of this image's [2,H,W] block in the gravity_original blob
  long long l_off;  // float offset of this image's [H,W] block in the latitude_original blob
  long long pix0;   // first global output-pixel index of this image (prefix sum of H*W)
};

// grid = (blocks over the largest image, n images); one thread = 4 consecutive output pixels of one row (16 B stores to each
// of the three planes).  Reads hit L1/L2 (the 320x320 source of an image is 1.2 MB); the kernel is bound by its writes.
__global__ void __launch_bounds__(256) postprocess_kernel(const float* __restrict__ vec, const float* __restrict__ lat, const PostImage* __restrict__ imgs, int n,
                                                          long long total, float* __restrict__ g_out, float* __restrict__ l_out, int lat_is_sin) {
  const int img = blockIdx.y;
  const PostImage im = imgs[img];
  const int W4 = (im.W + 3) >> 2;
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= im.H * W4) return;
  const int y = q / W4, x0 = (q - y * W4) * 4;
  const float sch = (float)kNet / (float)im.H, scw = (float)kNet / (float)im.W;
  const float sy = fmaxf(sch * ((float)y + 0.5f) - 0.5f, 0.f);
  const int y0 = min((int)sy, kNet - 1);
  const int y1 = y0 + (y0 < kNet - 1);
  const float ly = sy - (float)y0, hy = 1.f - ly;
  const float* v0 = vec + (long long)img * 2 * kNet * kNet;
  const float* v1 = v0 + kNet * kNet;
  const float* lp = lat + (long long)img * kNet * kNet;
  // the reference scales the field before resampling: vec * [[W/320],[H/320]] (float32 tensor built from python doubles)
  const float fx = (float)((double)im.W / (double)kNet), fy = (float)((double)im.H / (double)kNet);
  float ogx[4], ogy[4], ol[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int x = min(x0 + j, im.W - 1);
    const float sx = fmaxf(scw * ((float)x + 0.5f) - 0.5f, 0.f);
    const int xa = min((int)sx, kNet - 1);
    const int xb = xa + (xa < kNet - 1);
    const float lx = sx - (float)xa, hx = 1.f - lx;
    const int i00 = y0 * kNet + xa, i01 = y0 * kNet + xb, i10 = y1 * kNet + xa, i11 = y1 * kNet + xb;
    const float gx = hy * (hx * (v0[i00] * fx) + lx * (v0[i01] * fx)) + ly * (hx * (v0[i10] * fx) + lx * (v0[i11] * fx));
    const float gy = hy * (hx * (v1[i00] * fy) + lx * (v1[i01] * fy)) + ly * (hx * (v1[i10] * fy) + lx * (v1[i11] * fy));
    const float nrm = fmaxf(sqrtf(gx * gx + gy * gy), 1e-12f);
    ogx[j] = gx / nrm; ogy[j] = gy / nrm;
    float lv = hy * (hx * lp[i00] + lx * lp[i01]) + ly * (hx * lp[i10] + lx * lp[i11]);
    if (lat_is_sin) lv = asinf(lv) * (180.0f / 3.14159265358979323846f);
    ol[j] = lv;
  }
  const long long HW = (long long)im.H * im.W;
  const long long p = (long long)y * im.W + x0;
  float* gp = g_out + im.g_off + p;
  float* lpo = l_out + im.l_off + p;
  if (x0 + 3 < im.W && ((im.g_off + p) & 3) == 0 && ((im.g_off + HW + p) & 3) == 0 && ((im.l_off + p) & 3) == 0) {
    *reinterpret_cast<float4*>(gp) = make_float4(ogx[0], ogx[1], ogx[2], ogx[3]);
    *reinterpret_cast<float4*>(gp + HW) = make_float4(ogy[0], ogy[1], ogy[2], ogy[3]);
    *reinterpret_cast<float4*>(lpo) = make_float4(ol[0], ol[1], ol[2], ol[3]);
  } else {
    for (int j = 0; j < 4 && x0 + j < im.W; ++j) { gp[j] = ogx[j]; gp[HW + j] = ogy[j]; lpo[j] = ol[j]; }
  }
  (void)n; (void)total;
}

// =====================================================================================================
// Camera parameters -> dense perspective fields: PanoCam.get_up_general / get_lat_general
// (utils/panocam.py:451-556), what callers evaluate right after the inference path (utils/utils.py:367-385).
// One thread per pixel, float64 arithmetic like the numpy reference, float32 results; the kernel is bound by its stores.
//   up  [H, W, 2] (x, y): unit vector from the pixel centre (j + .5, i + .5) to the vertical vanishing point, flipped by
//                sign(elevation); the constant (-sin roll, -cos roll) when elevation == 0 exactly (:488)
//   lat [H, W] degrees: ray ((dx, dy, f) / f) rotated by roll, then elevation; -atan2(y_w, hypot(x_w, z_w)); dx / dy sample
//                linspace(-cx, W - cx, W) INCLUDING both end points (:534-539: spacing W / (W - 1), not pixel centres)
struct CamImage {
  int H, W;
  double f, cx, cy;           // focal length in pixels, principal point in pixels
  double sr, cr, se, ce;      // sin / cos of roll and elevation (computed on the host in float64)
  double sgn;                 // sign(elevation): +1, -1 or 0 (0 selects the constant field)
  long long up_off, lat_off;  // float offsets of this image's blocks in the output blobs
};
constexpr int kCamChunk = 24;   // images per launch (the descriptors travel as a kernel parameter)
struct CamBatch { CamImage im[kCamChunk]; };

__global__ void __launch_bounds__(256) camera_fields_kernel(const __grid_constant__ CamBatch batch, float* __restrict__ up, float* __restrict__ lat) {
  const CamImage& c = batch.im[blockIdx.y];
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= (long long)c.H * c.W) return;
  const int i = (int)(p / c.W), j = (int)(p - (long long)i * c.W);
  if (up) {
    double vx, vy;
    if (c.sgn == 0.0) { vx = -c.sr; vy = -c.cr; }
    else {
      const double vvp_x = (c.sr * c.ce * c.f) / -c.se + c.cx, vvp_y = (c.cr * c.ce * c.f) / -c.se + c.cy;
      vx = (vvp_x - ((double)j + 0.5)) * c.sgn;
      vy = (vvp_y - ((double)i + 0.5)) * c.sgn;
    }
    const double n = sqrt(vx * vx + vy * vy);
    *reinterpret_cast<float2*>(up + c.up_off + 2 * p) = make_float2((float)(vx / n), (float)(vy / n));
  }
  if (lat) {
    // numpy.linspace(start, stop, num): start + k * ((stop - start) / (num - 1)), last sample = stop exactly
    const double x0 = (-c.W / 2.0) - (c.cx - (c.W / 2.0)), x1 = (c.W / 2.0) - (c.cx - (c.W / 2.0));
    const double y0 = (-c.H / 2.0) - (c.cy - (c.H / 2.0)), y1 = (c.H / 2.0) - (c.cy - (c.H / 2.0));
    const double dx = c.W == 1 ? x0 : (j == c.W - 1 ? x1 : (double)j * ((x1 - x0) / (double)(c.W - 1)) + x0);
    const double dy = c.H == 1 ? y0 : (i == c.H - 1 ? y1 : (double)i * ((y1 - y0) / (double)(c.H - 1)) + y0);
    const double x = dx / c.f, y = dy / c.f;
    const double xw = x * c.cr - y * c.sr;
    const double yw = x * c.ce * c.sr + y * c.ce * c.cr - c.se;
    const double zw = x * c.se * c.sr + y * c.se * c.cr + c.ce;
    lat[c.lat_off + p] = (float)(-atan2(yw, sqrt(xw * xw + zw * zw)) / 3.14159265358979323846 * 180.0);
  }
}

}  // namespace pf
