/* rt_image.c — image containers, PNG decode (zlib inflate), PNG / QOI / PPM
 * encode, the default camera and the procedural environment map.
 *
 * Reference call sites: load_texture driver.c:106-116, output encode
 * driver.c:839-873 (Codin's PNG writer emits stored deflate blocks — the
 * committed output.png's IDAT is 1024*(1+3072)+overhead bytes), default camera
 * driver.c:765-767.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#include "rt_host.h"

bool rt_jpeg_decode(u8 const *bytes, size_t len, Image *out);

static _Thread_local char g_error[256];
void rt_host_set_error(char const *msg) { snprintf(g_error, sizeof g_error, "%s", msg); }
char const *rt_host_last_error(void) { return g_error; }

/* Buffers the GPU library DMA-reads (texels, BVH nodes, triangle block) and writes (the output image) come from
 * one hook, so a host can hand them out of pinned memory (rt_gpu_host_alloc: H2D then needs no staging copy).
 * A 64-byte header in front of each buffer remembers which release function belongs to it, so the hook may be
 * installed or changed at any time. */
typedef struct { void (*release)(void *); u8 pad[64 - sizeof(void (*)(void *))]; } Buffer_Header;
static void *(*g_buffer_alloc)(size_t) = NULL;
static void (*g_buffer_release)(void *) = NULL;

void rt_host_set_buffer_allocator(void *(*alloc)(size_t), void (*release)(void *)) {
  g_buffer_alloc = alloc;
  g_buffer_release = release;
}

void *rt_host_buffer_alloc(size_t bytes) {
  size_t total = ((bytes + 63) & ~(size_t)63) + sizeof(Buffer_Header);
  Buffer_Header *h = NULL;
  void (*release)(void *) = free;
  if (g_buffer_alloc && g_buffer_release) { h = g_buffer_alloc(total); release = g_buffer_release; }
  if (!h) { h = aligned_alloc(64, total); release = free; }
  if (!h) return NULL;
  h->release = release;
  return (u8 *)h + sizeof(Buffer_Header);
}

void rt_host_buffer_free(void *p) {
  if (!p) return;
  Buffer_Header *h = (Buffer_Header *)((u8 *)p - sizeof(Buffer_Header));
  h->release(h);
}

Image rt_image_alloc(isize width, isize height, i32 components) {
  Image im;
  memset(&im, 0, sizeof im);
  im.width = width; im.height = height; im.stride = width;
  im.components = components; im.pixel_type = PT_u8;
  im.pixels.len = width * height * components;
  im.pixels.data = rt_host_buffer_alloc((size_t)im.pixels.len);
  if (!im.pixels.data) { rt_host_set_error("out of memory"); memset(&im, 0, sizeof im); return im; }
  memset(im.pixels.data, 0, (size_t)im.pixels.len);
  return im;
}

void rt_image_free(Image *image) {
  rt_host_buffer_free(image->pixels.data);
  memset(image, 0, sizeof *image);
}

/* ------------------------------------------------------------------ PNG in */
static u32 be32(u8 const *p) { return ((u32)p[0] << 24) | ((u32)p[1] << 16) | ((u32)p[2] << 8) | p[3]; }

static int paeth(int a, int b, int c) {
  int p = a + b - c, pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
  return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

static bool png_decode(u8 const *bytes, size_t len, Image *out) {
  static const u8 sig[8] = { 0x89, 'P', 'N', 'G', '\r', '\n', 0x1a, '\n' };
  if (len < 8 || memcmp(bytes, sig, 8)) return false;
  u8 const *p = bytes + 8, *end = bytes + len;
  u32 w = 0, h = 0; int depth = 0, ctype = 0, interlace = 0;
  u8 *idat = NULL; size_t idat_len = 0;
  u8 palette[256 * 3]; memset(palette, 0, sizeof palette);
  bool ok = false;
  while (p + 12 <= end) {
    u32 n = be32(p);
    u8 const *type = p + 4, *body = p + 8;
    if (body + n + 4 > end) break;
    if (!memcmp(type, "IHDR", 4)) { w = be32(body); h = be32(body + 4); depth = body[8]; ctype = body[9]; interlace = body[12]; }
    else if (!memcmp(type, "PLTE", 4)) memcpy(palette, body, n < sizeof palette ? n : sizeof palette);
    else if (!memcmp(type, "IDAT", 4)) { idat = realloc(idat, idat_len + n); memcpy(idat + idat_len, body, n); idat_len += n; }
    else if (!memcmp(type, "IEND", 4)) break;
    p = body + n + 4;
  }
  int channels = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : 4;
  if (!w || !h || !idat || interlace || (depth != 8 && depth != 16)) {
    rt_host_set_error("png: unsupported (interlaced, or bit depth not 8/16)");
    free(idat);
    return false;
  }
  int bpp = channels * depth / 8;
  size_t row_bytes = (size_t)w * (size_t)bpp, raw_len = (row_bytes + 1) * h;
  u8 *raw = malloc(raw_len);
  uLongf got = (uLongf)raw_len;
  if (uncompress(raw, &got, idat, (uLong)idat_len) != Z_OK || got != raw_len) { rt_host_set_error("png: inflate failed"); goto done; }

  for (u32 y = 0; y < h; y++) {
    u8 *row = raw + y * (row_bytes + 1) + 1, *up = y ? row - (row_bytes + 1) : NULL;
    int filter = row[-1];
    for (size_t i = 0; i < row_bytes; i++) {
      int a = i >= (size_t)bpp ? row[i - bpp] : 0, b = up ? up[i] : 0, c = (up && i >= (size_t)bpp) ? up[i - bpp] : 0;
      int v = row[i];
      switch (filter) {
        case 1: v += a; break;
        case 2: v += b; break;
        case 3: v += (a + b) >> 1; break;
        case 4: v += paeth(a, b, c); break;
        default: break;
      }
      row[i] = (u8)v;
    }
  }
  /* stb keeps the file's channel count for RGB / RGBA and expands palettes;
   * grey becomes RGB here so the sampler's 3-channel reads stay in bounds. */
  {
    int out_ch = (ctype == 6 || ctype == 4) ? 4 : 3;
    *out = rt_image_alloc((isize)w, (isize)h, out_ch);
    int step = depth / 8;
    for (u32 y = 0; y < h; y++) {
      u8 const *row = raw + y * (row_bytes + 1) + 1;
      u8 *dst = out->pixels.data + (size_t)y * w * (size_t)out_ch;
      for (u32 x = 0; x < w; x++) {
        u8 const *s = row + (size_t)x * (size_t)bpp;
        u8 r, g, b, a = 255;
        switch (ctype) {
          case 0: r = g = b = s[0]; break;
          case 2: r = s[0]; g = s[step]; b = s[2 * step]; break;
          case 3: r = palette[3 * s[0]]; g = palette[3 * s[0] + 1]; b = palette[3 * s[0] + 2]; break;
          case 4: r = g = b = s[0]; a = s[step]; break;
          default: r = s[0]; g = s[step]; b = s[2 * step]; a = s[3 * step]; break;
        }
        dst[0] = r; dst[1] = g; dst[2] = b;
        if (out_ch == 4) dst[3] = a;
        dst += out_ch;
      }
    }
  }
  ok = true;
done:
  free(raw);
  free(idat);
  return ok;
}

/* With deferral on, a JPEG is NOT decoded here: the Image carries the compressed bytes (PT_RT_JPEG_BYTES) and the GPU
 * library decodes them during the scene upload.  Size comes from the first SOF0/SOF1 marker. */
static bool g_defer_jpeg = false;
void rt_host_defer_jpeg_decode(bool on) { g_defer_jpeg = on; }

static bool jpeg_frame_size(u8 const *p, size_t len, isize *w, isize *h, int *n_comp) {
  size_t i = 2;
  while (i + 4 <= len) {
    if (p[i] != 0xff) return false;
    u8 m = p[i + 1];
    if (m == 0xff) { i++; continue; }
    if (m == 0xd8 || (m >= 0xd0 && m <= 0xd7) || m == 0x01) { i += 2; continue; }
    size_t seg = ((size_t)p[i + 2] << 8) | p[i + 3];
    if (seg < 2 || i + 2 + seg > len) return false;
    if (m == 0xc0 || m == 0xc1) {
      if (seg < 8) return false;
      *h = ((isize)p[i + 5] << 8) | p[i + 6];
      *w = ((isize)p[i + 7] << 8) | p[i + 8];
      *n_comp = p[i + 9];
      return *w > 0 && *h > 0;
    }
    if (m == 0xc2 || m == 0xda) return false;        /* progressive, or a scan before any baseline frame */
    i += 2 + seg;
  }
  return false;
}

bool rt_image_decode(u8 const *bytes, size_t len, Image *out) {
  if (g_defer_jpeg && len >= 2 && bytes[0] == 0xff && bytes[1] == 0xd8) {
    isize w = 0, h = 0; int nc = 0;
    if (jpeg_frame_size(bytes, len, &w, &h, &nc) && (nc == 1 || nc == 3)) {
      memset(out, 0, sizeof *out);
      out->pixels.data = rt_host_buffer_alloc(len);
      if (!out->pixels.data) { rt_host_set_error("out of memory"); return false; }
      memcpy(out->pixels.data, bytes, len);
      out->pixels.len = (isize)len;
      out->width = w; out->height = h; out->stride = w;
      out->components = 3; out->pixel_type = PT_RT_JPEG_BYTES;
      return true;
    }
  }
  if (len >= 2 && bytes[0] == 0xff && bytes[1] == 0xd8) return rt_jpeg_decode(bytes, len, out);
  if (len >= 8 && bytes[0] == 0x89 && bytes[1] == 'P') return png_decode(bytes, len, out);
  rt_host_set_error("image: unknown format (baseline JPEG and PNG are supported)");
  return false;
}

static u8 *read_file(char const *path, size_t *len) {
  FILE *f = fopen(path, "rb");
  if (!f) return NULL;
  fseek(f, 0, SEEK_END);
  long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  u8 *buf = malloc((size_t)n + 1);
  if (fread(buf, 1, (size_t)n, f) != (size_t)n) { free(buf); fclose(f); return NULL; }
  fclose(f);
  buf[n] = 0;
  *len = (size_t)n;
  return buf;
}

bool rt_load_texture(char const *path, Image *out) {
  size_t len;
  u8 *bytes = read_file(path, &len);
  if (!bytes) { rt_host_set_error("texture: cannot read file"); return false; }
  bool ok = rt_image_decode(bytes, len, out);
  free(bytes);
  return ok;
}

/* ----------------------------------------------------------------- encoders */
static void put_be32(u8 *p, u32 v) { p[0] = (u8)(v >> 24); p[1] = (u8)(v >> 16); p[2] = (u8)(v >> 8); p[3] = (u8)v; }

static void png_chunk(FILE *f, char const *type, u8 const *body, u32 n) {
  u8 hdr[8];
  put_be32(hdr, n);
  memcpy(hdr + 4, type, 4);
  fwrite(hdr, 1, 8, f);
  if (n) fwrite(body, 1, n, f);
  uLong crc = crc32(0L, hdr + 4, 4);
  if (n) crc = crc32(crc, body, n);
  u8 tail[4];
  put_be32(tail, (u32)crc);
  fwrite(tail, 1, 4, f);
}

/* Stored (uncompressed) deflate, one filter-0 scanline stream. */
bool rt_save_png(char const *path, Image const *im) {
  if (im->components != 3 && im->components != 4) { rt_host_set_error("png: components"); return false; }
  FILE *f = fopen(path, "wb");
  if (!f) { rt_host_set_error("png: cannot open output"); return false; }
  static const u8 sig[8] = { 0x89, 'P', 'N', 'G', '\r', '\n', 0x1a, '\n' };
  fwrite(sig, 1, 8, f);
  u8 ihdr[13];
  put_be32(ihdr, (u32)im->width); put_be32(ihdr + 4, (u32)im->height);
  ihdr[8] = 8; ihdr[9] = im->components == 3 ? 2 : 6; ihdr[10] = ihdr[11] = ihdr[12] = 0;
  png_chunk(f, "IHDR", ihdr, 13);

  size_t row = (size_t)im->width * (size_t)im->components + 1, raw_len = row * (size_t)im->height;
  u8 *raw = malloc(raw_len);
  for (isize y = 0; y < im->height; y++) {
    raw[(size_t)y * row] = 0;
    memcpy(raw + (size_t)y * row + 1, im->pixels.data + (size_t)y * (size_t)im->stride * (size_t)im->components, row - 1);
  }
  size_t n_blocks = (raw_len + 65534) / 65535;
  size_t z_len = 2 + raw_len + 5 * n_blocks + 4;
  u8 *z = malloc(z_len), *w = z;
  *w++ = 0x78; *w++ = 0x01;
  for (size_t off = 0; off < raw_len; off += 65535) {
    size_t n = raw_len - off < 65535 ? raw_len - off : 65535;
    *w++ = (off + n == raw_len) ? 1 : 0;
    *w++ = (u8)n; *w++ = (u8)(n >> 8); *w++ = (u8)~n; *w++ = (u8)(~n >> 8);
    memcpy(w, raw + off, n);
    w += n;
  }
  put_be32(w, (u32)adler32(1L, raw, (uInt)raw_len));
  w += 4;
  png_chunk(f, "IDAT", z, (u32)(w - z));
  png_chunk(f, "IEND", NULL, 0);
  free(z);
  free(raw);
  fclose(f);
  return true;
}

bool rt_save_ppm(char const *path, Image const *im) {
  FILE *f = fopen(path, "wb");
  if (!f) { rt_host_set_error("ppm: cannot open output"); return false; }
  fprintf(f, "P6\n%ld %ld\n255\n", (long)im->width, (long)im->height);
  for (isize y = 0; y < im->height; y++)
    for (isize x = 0; x < im->width; x++)
      fwrite(im->pixels.data + (size_t)(x + y * im->stride) * (size_t)im->components, 1, 3, f);
  fclose(f);
  return true;
}

/* QOI 1.0 encoder (public spec: index / diff / luma / run / rgb / rgba ops). */
bool rt_save_qoi(char const *path, Image const *im) {
  FILE *f = fopen(path, "wb");
  if (!f) { rt_host_set_error("qoi: cannot open output"); return false; }
  int ch = im->components >= 4 ? 4 : 3;
  size_t cap = 14 + (size_t)im->width * (size_t)im->height * (size_t)(ch + 1) + 8;
  u8 *buf = malloc(cap), *w = buf;
  memcpy(w, "qoif", 4); w += 4;
  put_be32(w, (u32)im->width); w += 4;
  put_be32(w, (u32)im->height); w += 4;
  *w++ = (u8)ch; *w++ = 0;
  u8 index[64][4]; memset(index, 0, sizeof index);
  u8 prev[4] = { 0, 0, 0, 255 };
  int run = 0;
  isize total = im->width * im->height;
  for (isize i = 0; i < total; i++) {
    isize x = i % im->width, y = i / im->width;
    u8 const *s = im->pixels.data + (size_t)(x + y * im->stride) * (size_t)im->components;
    u8 px[4] = { s[0], s[1], s[2], (u8)(ch == 4 ? s[3] : 255) };
    if (!memcmp(px, prev, 4)) {
      run++;
      if (run == 62 || i == total - 1) { *w++ = (u8)(0xc0 | (run - 1)); run = 0; }
      continue;
    }
    if (run) { *w++ = (u8)(0xc0 | (run - 1)); run = 0; }
    int slot = (px[0] * 3 + px[1] * 5 + px[2] * 7 + px[3] * 11) % 64;
    if (!memcmp(index[slot], px, 4)) {
      *w++ = (u8)slot;
    } else {
      memcpy(index[slot], px, 4);
      if (px[3] == prev[3]) {
        signed char dr = (signed char)(px[0] - prev[0]), dg = (signed char)(px[1] - prev[1]), db = (signed char)(px[2] - prev[2]);
        signed char dr_dg = (signed char)(dr - dg), db_dg = (signed char)(db - dg);
        if (dr > -3 && dr < 2 && dg > -3 && dg < 2 && db > -3 && db < 2) {
          *w++ = (u8)(0x40 | ((dr + 2) << 4) | ((dg + 2) << 2) | (db + 2));
        } else if (dr_dg > -9 && dr_dg < 8 && dg > -33 && dg < 32 && db_dg > -9 && db_dg < 8) {
          *w++ = (u8)(0x80 | (dg + 32));
          *w++ = (u8)(((dr_dg + 8) << 4) | (db_dg + 8));
        } else {
          *w++ = 0xfe; *w++ = px[0]; *w++ = px[1]; *w++ = px[2];
        }
      } else {
        *w++ = 0xff; *w++ = px[0]; *w++ = px[1]; *w++ = px[2]; *w++ = px[3];
      }
    }
    memcpy(prev, px, 4);
  }
  static const u8 tail[8] = { 0, 0, 0, 0, 0, 0, 0, 1 };
  memcpy(w, tail, 8); w += 8;
  fwrite(buf, 1, (size_t)(w - buf), f);
  free(buf);
  fclose(f);
  return true;
}

static bool has_suffix(char const *s, char const *suffix) {
  size_t n = strlen(s), m = strlen(suffix);
  return n >= m && !strcmp(s + n - m, suffix);
}

bool rt_save_image(char const *path, Image const *image) {
  if (has_suffix(path, ".qoi")) return rt_save_qoi(path, image);
  if (has_suffix(path, ".ppm")) return rt_save_ppm(path, image);
  if (!has_suffix(path, ".png"))
    fprintf(stdout, "output format not recognized for output path '%s', defaulting to png\n", path);
  return rt_save_png(path, image);
}

/* ------------------------------------------------------------------- camera */
void rt_camera_default(Camera *camera) {
  memset(camera, 0, sizeof *camera);
  for (int i = 0; i < 4; i++) camera->view_matrix.rows[i][i] = 1.0f;
  camera->view_matrix.rows[2][3] = 3.0f;
  camera->fov          = (f32)((70.0f / 360.0f) * 3.14159265358979323846 * 2.0);
  camera->focal_length = 1.0f / tanf(camera->fov * 0.5f);
}

void rt_camera_look_at(Camera *camera, Vec3 eye, Vec3 target, Vec3 up, f32 fov_radians) {
  f32 f[3] = { target.x - eye.x, target.y - eye.y, target.z - eye.z };
  f32 fl = sqrtf(f[0] * f[0] + f[1] * f[1] + f[2] * f[2]);
  for (int i = 0; i < 3; i++) f[i] /= fl;
  f32 r[3] = { f[1] * up.z - f[2] * up.y, f[2] * up.x - f[0] * up.z, f[0] * up.y - f[1] * up.x };
  f32 rl = sqrtf(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
  for (int i = 0; i < 3; i++) r[i] /= rl;
  f32 u[3] = { r[1] * f[2] - r[2] * f[1], r[2] * f[0] - r[0] * f[2], r[0] * f[1] - r[1] * f[0] };
  memset(camera, 0, sizeof *camera);
  /* camera looks down its -Z: columns are right, up, -forward, eye */
  for (int i = 0; i < 3; i++) {
    camera->view_matrix.rows[i][0] = r[i];
    camera->view_matrix.rows[i][1] = u[i];
    camera->view_matrix.rows[i][2] = -f[i];
  }
  camera->view_matrix.rows[0][3] = eye.x;
  camera->view_matrix.rows[1][3] = eye.y;
  camera->view_matrix.rows[2][3] = eye.z;
  camera->view_matrix.rows[3][3] = 1.0f;
  camera->fov = fov_radians;
  camera->focal_length = 1.0f / tanf(fov_radians * 0.5f);
}

/* -------------------------------------------------- procedural environment */
static int tri_wave(int v, int period) {       /* 0..period/2..0 */
  int m = v % period;
  return m < period / 2 ? m : period - m;
}

void rt_generate_background(Image *out, isize width, isize height) {
  *out = rt_image_alloc(width, height, 3);
  isize sun_x = width * 5 / 8, sun_y = height * 5 / 16, sun_r = height / 24;
  for (isize y = 0; y < height; y++) {
    for (isize x = 0; x < width; x++) {
      u8 *p = out->pixels.data + 3 * (x + y * width);
      int r, g, b;
      isize horizon = height / 2;
      if (y < horizon) {
        /* zenith (58,104,196) -> horizon (214,226,240), quadratic in elevation */
        int t = (int)(y * 1024 / horizon);
        t = t * t / 1024;
        r = 58 + (214 - 58) * t / 1024;
        g = 104 + (226 - 104) * t / 1024;
        b = 196 + (240 - 196) * t / 1024;
        /* soft cloud bands */
        int band = tri_wave((int)(x * 3 + y * 5), (int)(width / 4 + 1)) * 2048 / (int)(width / 4 + 1);
        int lift = band * tri_wave((int)y * 7, (int)(height / 3 + 1)) / (int)(height / 3 + 1);
        r += lift * 24 / 1024; g += lift * 20 / 1024; b += lift * 10 / 1024;
      } else {
        /* ground: horizon haze (150,140,128) -> nadir (46,40,34), with a tiled pattern */
        int t = (int)((y - horizon) * 1024 / (height - horizon));
        r = 150 + (46 - 150) * t / 1024;
        g = 140 + (40 - 140) * t / 1024;
        b = 128 + (34 - 128) * t / 1024;
        int cell = (int)((x * 32 / width) + ((y - horizon) * 16 / (height - horizon)));
        if (cell & 1) { r = r * 7 / 8; g = g * 7 / 8; b = b * 7 / 8; }
      }
      isize dx = x - sun_x, dy = y - sun_y;
      isize d2 = dx * dx + dy * dy, r2 = sun_r * sun_r;
      if (d2 < r2) { r = 255; g = 250; b = 236; }
      else if (d2 < 16 * r2) {
        int glow = (int)((16 * r2 - d2) * 96 / (15 * r2));
        r += glow; g += glow * 9 / 10; b += glow * 6 / 10;
      }
      p[0] = (u8)(r > 255 ? 255 : r);
      p[1] = (u8)(g > 255 ? 255 : g);
      p[2] = (u8)(b > 255 ? 255 : b);
    }
  }
}
