/* rt_jpeg.c — baseline (SOF0) Huffman JPEG decoder.
 *
 * The reference decodes glTF-embedded textures with stb_image through Codin
 * (reference driver.c:620-626); neither is in the tree.  This decoder follows
 * the published stb_image pipeline so decoded texels are as close to the
 * reference's as can be arranged: the "slow integer" 8x8 IDCT with 12-bit
 * constants (column pass >>10, row pass >>17 with +128 bias), the 3:1 / 9:3:3:1
 * triangle-filter chroma upsampling for 2x1 / 2x2 subsampling, and fixed-point
 * BT.601 YCbCr->RGB with 20 fractional bits.  UNPINNED: bit-equality with the
 * stb build the author used cannot be checked here; tests compare against
 * PIL/libjpeg within a small mean error.
 */
#include <stdlib.h>
#include <string.h>

#include "rt_host.h"

void rt_host_set_error(char const *msg);

typedef struct {
  u8  fast_len[512];     /* 9-bit lookahead: code length or 0 */
  u8  fast_sym[512];
  u16 code[256];
  u8  size[257];
  u8  values[256];
  u32 maxcode[18];
  int delta[17];
} Huff;

typedef struct {
  int id, h, v, tq, td, ta;
  int dc_pred;
  int x, y, w2, h2;       /* real and MCU-padded sample dimensions */
  u8 *data;
} Comp;

typedef struct {
  u8 const *p, *end;
  u32 bits; int n_bits;
  int marker;             /* pending marker hit while filling, or 0 */
  int eob_run;
  Huff dc[4], ac[4];
  u16 dequant[4][64];
  Comp comp[4];
  int n_comp, width, height, h_max, v_max, mcu_w, mcu_h, mcus_x, mcus_y;
  int restart_interval, todo;
} Dec;

static const u8 zigzag[64 + 15] = {
   0,  1,  8, 16,  9,  2,  3, 10, 17, 24, 32, 25, 18, 11,  4,  5,
  12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13,  6,  7, 14, 21, 28,
  35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51,
  58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63,
  63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63 };

static bool build_huff(Huff *h, u8 const *counts) {
  int k = 0;
  for (int len = 1; len <= 16; len++)
    for (int i = 0; i < counts[len - 1]; i++) { if (k >= 256) return false; h->size[k++] = (u8)len; }
  h->size[k] = 0;
  u32 code = 0;
  k = 0;
  for (int len = 1; len <= 16; len++) {
    h->delta[len] = k - (int)code;
    if (h->size[k] == len) {
      while (h->size[k] == len) h->code[k++] = (u16)code++;
      if (code - 1 >= (1u << len)) return false;
    }
    h->maxcode[len] = code << (16 - len);
    code <<= 1;
  }
  h->maxcode[17] = 0xffffffffu;
  memset(h->fast_len, 0, sizeof h->fast_len);
  for (int i = 0; i < k; i++) {
    int s = h->size[i];
    if (s <= 9) {
      int c = h->code[i] << (9 - s), m = 1 << (9 - s);
      for (int j = 0; j < m; j++) { h->fast_len[c + j] = (u8)s; h->fast_sym[c + j] = (u8)i; }
    }
  }
  return true;
}

static void fill_bits(Dec *d) {
  while (d->n_bits <= 24) {
    int b = 0;
    if (!d->marker && d->p < d->end) {
      b = *d->p++;
      if (b == 0xff) {
        int c = d->p < d->end ? *d->p++ : 0xd9;
        while (c == 0xff) c = d->p < d->end ? *d->p++ : 0xd9;
        if (c != 0) { d->marker = c; b = 0; }
      }
    }
    d->bits |= (u32)b << (24 - d->n_bits);
    d->n_bits += 8;
  }
}

static int decode_symbol(Dec *d, Huff const *h) {
  if (d->n_bits < 16) fill_bits(d);
  int look = (int)(d->bits >> 23);
  int s = h->fast_len[look];
  if (s) {
    d->bits <<= s; d->n_bits -= s;
    return h->values[h->fast_sym[look]];
  }
  u32 top = d->bits >> 16;
  int len;
  for (len = 10; len <= 16; len++) if (top < h->maxcode[len]) break;
  if (len > 16) return -1;
  int idx = (int)((d->bits >> (32 - len)) & ((1u << len) - 1)) + h->delta[len];
  if (idx < 0 || idx > 255) return -1;
  d->bits <<= len; d->n_bits -= len;
  return h->values[idx];
}

/* n-bit signed magnitude (JPEG "extend") */
static int receive_extend(Dec *d, int n) {
  if (n == 0) return 0;
  if (d->n_bits < n) fill_bits(d);
  int v = (int)(d->bits >> (32 - n));
  d->bits <<= n; d->n_bits -= n;
  if (v < (1 << (n - 1))) v -= (1 << n) - 1;
  return v;
}

static bool decode_block(Dec *d, short blk[64], Comp *c) {
  memset(blk, 0, 64 * sizeof(short));
  Huff const *hdc = &d->dc[c->td], *hac = &d->ac[c->ta];
  u16 const *dq = d->dequant[c->tq];
  int t = decode_symbol(d, hdc);
  if (t < 0 || t > 15) return false;
  int diff = receive_extend(d, t);
  c->dc_pred += diff;
  blk[0] = (short)(c->dc_pred * dq[0]);
  for (int k = 1; k < 64;) {
    int rs = decode_symbol(d, hac);
    if (rs < 0) return false;
    int s = rs & 15, r = rs >> 4;
    if (s == 0) {
      if (rs != 0xf0) break;
      k += 16;
    } else {
      k += r;
      int z = zigzag[k++];
      blk[z] = (short)(receive_extend(d, s) * dq[z]);
    }
  }
  return true;
}

static inline u8 clamp_u8(int x) { return (u8)(x < 0 ? 0 : (x > 255 ? 255 : x)); }

#define F2F(x) ((int)((x) * 4096 + 0.5))
#define FSH(x) ((x) * 4096)

#define IDCT_1D(s0, s1, s2, s3, s4, s5, s6, s7)                 \
  int t0, t1, t2, t3, p1, p2, p3, p4, p5, x0, x1, x2, x3;       \
  p2 = s2; p3 = s6;                                             \
  p1 = (p2 + p3) * F2F(0.5411961f);                             \
  t2 = p1 + p3 * F2F(-1.847759065f);                            \
  t3 = p1 + p2 * F2F(0.765366865f);                             \
  p2 = s0; p3 = s4;                                             \
  t0 = FSH(p2 + p3); t1 = FSH(p2 - p3);                         \
  x0 = t0 + t3; x3 = t0 - t3; x1 = t1 + t2; x2 = t1 - t2;       \
  t0 = s7; t1 = s5; t2 = s3; t3 = s1;                           \
  p3 = t0 + t2; p4 = t1 + t3; p1 = t0 + t3; p2 = t1 + t2;       \
  p5 = (p3 + p4) * F2F(1.175875602f);                           \
  t0 = t0 * F2F(0.298631336f);                                  \
  t1 = t1 * F2F(2.053119869f);                                  \
  t2 = t2 * F2F(3.072711026f);                                  \
  t3 = t3 * F2F(1.501321110f);                                  \
  p1 = p5 + p1 * F2F(-0.899976223f);                            \
  p2 = p5 + p2 * F2F(-2.562915447f);                            \
  p3 = p3 * F2F(-1.961570560f);                                 \
  p4 = p4 * F2F(-0.390180644f);                                 \
  t3 += p1 + p4; t2 += p2 + p3; t1 += p2 + p4; t0 += p1 + p3;

static void idct_block(u8 *out, int stride, short const data[64]) {
  int val[64], *v = val;
  short const *dd = data;
  for (int i = 0; i < 8; i++, dd++, v++) {
    if (dd[8] == 0 && dd[16] == 0 && dd[24] == 0 && dd[32] == 0 && dd[40] == 0 && dd[48] == 0 && dd[56] == 0) {
      int dc = dd[0] * 4;
      v[0] = v[8] = v[16] = v[24] = v[32] = v[40] = v[48] = v[56] = dc;
    } else {
      IDCT_1D(dd[0], dd[8], dd[16], dd[24], dd[32], dd[40], dd[48], dd[56])
      x0 += 512; x1 += 512; x2 += 512; x3 += 512;
      v[0]  = (x0 + t3) >> 10; v[56] = (x0 - t3) >> 10;
      v[8]  = (x1 + t2) >> 10; v[48] = (x1 - t2) >> 10;
      v[16] = (x2 + t1) >> 10; v[40] = (x2 - t1) >> 10;
      v[24] = (x3 + t0) >> 10; v[32] = (x3 - t0) >> 10;
    }
  }
  v = val;
  for (int i = 0; i < 8; i++, v += 8, out += stride) {
    IDCT_1D(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7])
    x0 += 65536 + (128 << 17); x1 += 65536 + (128 << 17);
    x2 += 65536 + (128 << 17); x3 += 65536 + (128 << 17);
    out[0] = clamp_u8((x0 + t3) >> 17); out[7] = clamp_u8((x0 - t3) >> 17);
    out[1] = clamp_u8((x1 + t2) >> 17); out[6] = clamp_u8((x1 - t2) >> 17);
    out[2] = clamp_u8((x2 + t1) >> 17); out[5] = clamp_u8((x2 - t1) >> 17);
    out[3] = clamp_u8((x3 + t0) >> 17); out[4] = clamp_u8((x3 - t0) >> 17);
  }
}

static void reset_entropy(Dec *d) {
  d->bits = 0; d->n_bits = 0; d->marker = 0;
  for (int i = 0; i < d->n_comp; i++) d->comp[i].dc_pred = 0;
  d->todo = d->restart_interval ? d->restart_interval : 0x7fffffff;
}

static bool decode_scan(Dec *d, int n_scan, int const *order) {
  reset_entropy(d);
  short blk[64];
  if (n_scan == 1) {
    Comp *c = &d->comp[order[0]];
    int bw = (c->x + 7) >> 3, bh = (c->y + 7) >> 3;
    for (int j = 0; j < bh; j++) for (int i = 0; i < bw; i++) {
      if (!decode_block(d, blk, c)) return false;
      idct_block(c->data + c->w2 * j * 8 + i * 8, c->w2, blk);
      if (--d->todo <= 0) {
        if (d->n_bits < 24) fill_bits(d);
        if (!(d->marker >= 0xd0 && d->marker <= 0xd7)) return true;
        reset_entropy(d);
      }
    }
    return true;
  }
  for (int j = 0; j < d->mcus_y; j++) for (int i = 0; i < d->mcus_x; i++) {
    for (int k = 0; k < n_scan; k++) {
      Comp *c = &d->comp[order[k]];
      for (int y = 0; y < c->v; y++) for (int x = 0; x < c->h; x++) {
        int bx = (i * c->h + x) * 8, by = (j * c->v + y) * 8;
        if (!decode_block(d, blk, c)) return false;
        idct_block(c->data + c->w2 * by + bx, c->w2, blk);
      }
    }
    if (--d->todo <= 0) {
      if (d->n_bits < 24) fill_bits(d);
      if (!(d->marker >= 0xd0 && d->marker <= 0xd7)) return true;
      reset_entropy(d);
    }
  }
  return true;
}

/* ---- chroma upsampling, one output row at a time ---- */
static u8 *up_copy(u8 *out, u8 const *near, u8 const *far, int w, int hs) { (void)out; (void)far; (void)w; (void)hs; return (u8 *)near; }

static u8 *up_v2(u8 *out, u8 const *near, u8 const *far, int w, int hs) {
  (void)hs;
  for (int i = 0; i < w; i++) out[i] = (u8)((3 * near[i] + far[i] + 2) >> 2);
  return out;
}

static u8 *up_h2(u8 *out, u8 const *in, u8 const *far, int w, int hs) {
  (void)far; (void)hs;
  if (w == 1) { out[0] = out[1] = in[0]; return out; }
  out[0] = in[0];
  out[1] = (u8)((in[0] * 3 + in[1] + 2) >> 2);
  int i;
  for (i = 1; i < w - 1; i++) {
    int n = 3 * in[i] + 2;
    out[i * 2 + 0] = (u8)((n + in[i - 1]) >> 2);
    out[i * 2 + 1] = (u8)((n + in[i + 1]) >> 2);
  }
  out[i * 2 + 0] = (u8)((in[w - 2] * 3 + in[w - 1] + 2) >> 2);
  out[i * 2 + 1] = in[w - 1];
  return out;
}

static u8 *up_hv2(u8 *out, u8 const *near, u8 const *far, int w, int hs) {
  (void)hs;
  if (w == 1) { out[0] = out[1] = (u8)((3 * near[0] + far[0] + 2) >> 2); return out; }
  int t1 = 3 * near[0] + far[0], t0;
  out[0] = (u8)((t1 + 2) >> 2);
  for (int i = 1; i < w; i++) {
    t0 = t1;
    t1 = 3 * near[i] + far[i];
    out[i * 2 - 1] = (u8)((3 * t0 + t1 + 8) >> 4);
    out[i * 2]     = (u8)((3 * t1 + t0 + 8) >> 4);
  }
  out[w * 2 - 1] = (u8)((t1 + 2) >> 2);
  return out;
}

static u8 *up_generic(u8 *out, u8 const *near, u8 const *far, int w, int hs) {
  (void)far;
  for (int i = 0; i < w; i++) for (int j = 0; j < hs; j++) out[i * hs + j] = near[i];
  return out;
}

typedef u8 *(*Upsample)(u8 *, u8 const *, u8 const *, int, int);

#define FIX20(x) (((int)((x) * 4096.0f + 0.5f)) << 8)

static void ycc_row(u8 *out, u8 const *y, u8 const *cb, u8 const *cr, int n) {
  for (int i = 0; i < n; i++) {
    int yf = (y[i] << 20) + (1 << 19);
    int r_ = cr[i] - 128, b_ = cb[i] - 128;
    int r = yf + r_ * FIX20(1.40200f);
    int g = yf + (r_ * -FIX20(0.71414f)) + ((b_ * -FIX20(0.34414f)) & 0xffff0000);
    int b = yf + b_ * FIX20(1.77200f);
    r >>= 20; g >>= 20; b >>= 20;
    out[0] = clamp_u8(r); out[1] = clamp_u8(g); out[2] = clamp_u8(b);
    out += 3;
  }
}

static int be16(u8 const *p) { return (p[0] << 8) | p[1]; }

bool rt_jpeg_decode(u8 const *bytes, size_t len, Image *out) {
  Dec *d = calloc(1, sizeof *d);
  bool ok = false, seen_sof = false, done = false;
  u8 const *p = bytes, *end = bytes + len;
  if (len < 4 || p[0] != 0xff || p[1] != 0xd8) { rt_host_set_error("jpeg: no SOI"); goto out; }
  p += 2;
  while (!done && p + 4 <= end) {
    if (*p != 0xff) { p++; continue; }
    int m = p[1];
    p += 2;
    if (m == 0xff) { p--; continue; }
    if (m == 0xd9) break;
    if (m == 0x01 || (m >= 0xd0 && m <= 0xd7)) continue;
    int L = be16(p);
    u8 const *seg = p + 2, *seg_end = p + L;
    if (seg_end > end) { rt_host_set_error("jpeg: truncated segment"); goto out; }
    switch (m) {
      case 0xdb:
        while (seg < seg_end) {
          int pq = seg[0] >> 4, tq = seg[0] & 15;
          if (tq > 3) goto bad;
          seg++;
          for (int i = 0; i < 64; i++) { d->dequant[tq][zigzag[i]] = (u16)(pq ? be16(seg) : seg[0]); seg += pq ? 2 : 1; }
        }
        break;
      case 0xc4:
        while (seg < seg_end) {
          int tc = seg[0] >> 4, th = seg[0] & 15;
          if (tc > 1 || th > 3) goto bad;
          Huff *h = tc ? &d->ac[th] : &d->dc[th];
          int total = 0;
          for (int i = 0; i < 16; i++) total += seg[1 + i];
          if (total > 256 || !build_huff(h, seg + 1)) goto bad;
          memcpy(h->values, seg + 17, (size_t)total);
          seg += 17 + total;
        }
        break;
      case 0xc0: case 0xc1: {
        if (seg[0] != 8) { rt_host_set_error("jpeg: only 8-bit"); goto out; }
        d->height = be16(seg + 1); d->width = be16(seg + 3); d->n_comp = seg[5];
        if (d->n_comp != 1 && d->n_comp != 3) { rt_host_set_error("jpeg: component count"); goto out; }
        for (int i = 0; i < d->n_comp; i++) {
          Comp *c = &d->comp[i];
          c->id = seg[6 + 3 * i]; c->h = seg[7 + 3 * i] >> 4; c->v = seg[7 + 3 * i] & 15; c->tq = seg[8 + 3 * i];
          if (!c->h || !c->v || c->h > 4 || c->v > 4 || c->tq > 3) goto bad;
          if (c->h > d->h_max) d->h_max = c->h;
          if (c->v > d->v_max) d->v_max = c->v;
        }
        d->mcu_w = d->h_max * 8; d->mcu_h = d->v_max * 8;
        d->mcus_x = (d->width + d->mcu_w - 1) / d->mcu_w;
        d->mcus_y = (d->height + d->mcu_h - 1) / d->mcu_h;
        for (int i = 0; i < d->n_comp; i++) {
          Comp *c = &d->comp[i];
          c->x = (d->width * c->h + d->h_max - 1) / d->h_max;
          c->y = (d->height * c->v + d->v_max - 1) / d->v_max;
          c->w2 = d->mcus_x * c->h * 8;
          c->h2 = d->mcus_y * c->v * 8;
          c->data = malloc((size_t)c->w2 * (size_t)c->h2 + 15);
        }
        seen_sof = true;
      } break;
      case 0xc2: rt_host_set_error("jpeg: progressive not supported"); goto out;
      case 0xdd: d->restart_interval = be16(seg); break;
      case 0xda: {
        if (!seen_sof) goto bad;
        int ns = seg[0], order[4];
        if (ns < 1 || ns > d->n_comp) goto bad;
        for (int i = 0; i < ns; i++) {
          int id = seg[1 + 2 * i], which = -1;
          for (int k = 0; k < d->n_comp; k++) if (d->comp[k].id == id) which = k;
          if (which < 0) goto bad;
          d->comp[which].td = seg[2 + 2 * i] >> 4;
          d->comp[which].ta = seg[2 + 2 * i] & 15;
          if (d->comp[which].td > 3 || d->comp[which].ta > 3) goto bad;
          order[i] = which;
        }
        d->p = seg_end; d->end = end;
        if (!decode_scan(d, ns, order)) { rt_host_set_error("jpeg: corrupt scan"); goto out; }
        /* baseline: one interleaved scan, or one scan per component */
        if (ns < d->n_comp) { p = d->marker ? d->p - 2 : d->p; continue; }
        done = true;
      } break;
      default: break;
    }
    p = seg_end;
  }
  if (!seen_sof) { rt_host_set_error("jpeg: no frame"); goto out; }

  *out = rt_image_alloc(d->width, d->height, 3);
  {
    u8 *line[3] = {0};
    struct { Upsample fn; int hs, vs, ystep, ypos, w_lo; u8 const *l0, *l1; } r[3];
    for (int k = 0; k < d->n_comp; k++) {
      Comp *c = &d->comp[k];
      line[k] = malloc((size_t)d->width + 8 + (size_t)d->h_max * 8);
      r[k].hs = d->h_max / c->h; r[k].vs = d->v_max / c->v;
      r[k].ystep = r[k].vs >> 1;
      r[k].w_lo = (d->width + r[k].hs - 1) / r[k].hs;
      r[k].ypos = 0;
      r[k].l0 = r[k].l1 = c->data;
      if      (r[k].hs == 1 && r[k].vs == 1) r[k].fn = up_copy;
      else if (r[k].hs == 1 && r[k].vs == 2) r[k].fn = up_v2;
      else if (r[k].hs == 2 && r[k].vs == 1) r[k].fn = up_h2;
      else if (r[k].hs == 2 && r[k].vs == 2) r[k].fn = up_hv2;
      else                                   r[k].fn = up_generic;
    }
    for (int j = 0; j < d->height; j++) {
      u8 *rows[3];
      for (int k = 0; k < d->n_comp; k++) {
        int y_bot = r[k].ystep >= (r[k].vs >> 1);
        rows[k] = r[k].fn(line[k], y_bot ? r[k].l1 : r[k].l0, y_bot ? r[k].l0 : r[k].l1, r[k].w_lo, r[k].hs);
        if (++r[k].ystep >= r[k].vs) {
          r[k].ystep = 0;
          r[k].l0 = r[k].l1;
          if (++r[k].ypos < d->comp[k].y) r[k].l1 += d->comp[k].w2;
        }
      }
      u8 *dst = out->pixels.data + (size_t)j * (size_t)d->width * 3;
      if (d->n_comp == 3) ycc_row(dst, rows[0], rows[1], rows[2], d->width);
      else for (int i = 0; i < d->width; i++) dst[3 * i] = dst[3 * i + 1] = dst[3 * i + 2] = rows[0][i];
    }
    for (int k = 0; k < 3; k++) free(line[k]);
  }
  ok = true;
  goto out;
bad:
  rt_host_set_error("jpeg: malformed header");
out:
  for (int i = 0; i < 4; i++) free(d->comp[i].data);
  free(d);
  return ok;
}
