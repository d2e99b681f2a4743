/* TEST INFRASTRUCTURE ONLY -- plain-C restatement of torchvision.ops.roi_pool (forward, NCHW, fp32),
 * the un-vendored dependency the reference calls at detectron2/modeling/poolers.py:223-226
 * (torchvision 0.26.0 here; SURVEY.md §8c restates and verifies the semantics).  Built by
 * __graft_entry__.build() into oracle/_build/libroipool_ref.so and used by tests/ as the bit-exact
 * checker of drn_roipool_fwd at sizes where the numpy loop is too slow.  Never linked by the product. */
#include <math.h>
#include <float.h>

static int imin(int a, int b) { return a < b ? a : b; }
static int imax(int a, int b) { return a > b ? a : b; }

/* feat: [C][h][w]; boxes: [R][4] (x0,y0,x1,y1) image px; out: [R][C][P][P] */
void roipool_ref(const float* feat, int C, int h, int w, const float* boxes, int R, float scale, int P, float* out) {
  for (int r = 0; r < R; ++r) {
    const float* b = boxes + 4 * r;
    const int sw = (int)roundf(b[0] * scale), sh = (int)roundf(b[1] * scale);
    const int ew = (int)roundf(b[2] * scale), eh = (int)roundf(b[3] * scale);
    const int rw = imax(ew - sw + 1, 1), rh = imax(eh - sh + 1, 1);
    const float bh = (float)rh / (float)P, bw = (float)rw / (float)P;
    for (int ph = 0; ph < P; ++ph) {
      const int hs = imin(imax((int)floorf((float)ph * bh) + sh, 0), h);
      const int he = imin(imax((int)ceilf((float)(ph + 1) * bh) + sh, 0), h);
      for (int pw = 0; pw < P; ++pw) {
        const int ws = imin(imax((int)floorf((float)pw * bw) + sw, 0), w);
        const int we = imin(imax((int)ceilf((float)(pw + 1) * bw) + sw, 0), w);
        const int empty = (he <= hs) || (we <= ws);
        for (int c = 0; c < C; ++c) {
          float m = empty ? 0.f : -FLT_MAX;
          const float* f = feat + (long)c * h * w;
          for (int y = hs; y < he; ++y)
            for (int x = ws; x < we; ++x)
              if (f[y * w + x] > m) m = f[y * w + x];
          out[(((long)r * C + c) * P + ph) * P + pw] = m;
        }
      }
    }
  }
}
