/* oracle_denoise.c — TEST INFRASTRUCTURE.  Restatement of the reference's
 * 3x3 luminance-median denoiser (reference denoiser.c:9-149).  Works on the
 * u8 sRGB-encoded image, clamp-to-edge taps, stable insertion sort by
 * luminance, blend toward the median by a noise-adaptive factor.
 */
#include <pthread.h>
#include <stdatomic.h>
#include <stdlib.h>
#include <assert.h>

#include "oracle.h"
#include "oracle_vec.h"

#define THRESHOLD        0.0125f   /* denoiser.c:9  */
#define NEIGHBOUR_WEIGHT 5         /* denoiser.c:10 */

typedef struct { Vec3 rgb; f32 luma; } Tap;

/* denoiser.c:16-27 */
static inline Vec3 fetch(Image const *img, isize x, isize y) {
  if (x < 0) x = 0;
  if (y < 0) y = 0;
  if (x >= img->width)  x = img->width - 1;
  if (y >= img->height) y = img->height - 1;
  Vec3 c = v3(0, 0, 0);
  int n = img->components < 3 ? img->components : 3;
  for (int k = 0; k < n; k++) c.data[k] = img->pixels.data[(x + y * img->stride) * img->components + k] / 255.999f;
  return c;
}

static void filter_pixel(Image const *src, Image const *dst, isize x, isize y) {
  Tap sorted[9];
  Tap centre = {0};
  int count = 0;
  /* denoiser.c:80-107: insert before the first strictly greater luminance */
  for (int oy = -1; oy <= 1; oy++) {
    for (int ox = -1; ox <= 1; ox++) {
      Tap t;
      t.rgb  = fetch(src, x + ox, y + oy);
      t.luma = v3_dot(t.rgb, v3(0.2126f, 0.7152f, 0.0722f));
      if (ox == 0 && oy == 0) centre = t;
      int pos = count;
      for (int i = 0; i < count; i++) {
        if (sorted[i].luma > t.luma) { pos = i; break; }
      }
      for (int j = count; j > pos; j--) sorted[j] = sorted[j - 1];
      sorted[pos] = t;
      count++;
    }
  }
  /* denoiser.c:109-121 */
  Tap median = sorted[4];
  f32 mean = 0;
  for (int i = 1; i < 8; i++) mean += sorted[i].luma;
  mean /= 7;
  f32 noisiness = f32_abs(median.luma - mean);
  f32 diff = f32_abs(median.luma - centre.luma) - noisiness * NEIGHBOUR_WEIGHT;
  diff = f32_clamp(diff, 0, THRESHOLD) / THRESHOLD;
  Vec3 blended = v3_lerp(centre.rgb, median.rgb, diff);
  /* denoiser.c:29-38 */
  int n = dst->components < 3 ? dst->components : 3;
  for (int k = 0; k < n; k++) dst->pixels.data[(x + y * dst->stride) * dst->components + k] = (u8)(blended.data[k] * 255.999f);
}

typedef struct {
  Image const *src, *dst;
  atomic_long  next_chunk;
} Job;

/* denoiser.c:47-127 */
static void *worker(void *arg) {
  Job *job = arg;
  isize width = job->src->width, height = job->src->height;
  isize chunks_x = (width + RT_CHUNK_SIZE - 1) / RT_CHUNK_SIZE;
  isize chunks_y = (height + RT_CHUNK_SIZE - 1) / RT_CHUNK_SIZE;
  for (;;) {
    isize c = atomic_fetch_add(&job->next_chunk, 1);
    if (c >= chunks_x * chunks_y) break;
    isize x0 = (c % chunks_x) * RT_CHUNK_SIZE, y0 = (c / chunks_x) * RT_CHUNK_SIZE;
    for (isize y = y0; y < y0 + RT_CHUNK_SIZE && y < height; y++)
      for (isize x = x0; x < x0 + RT_CHUNK_SIZE && x < width; x++)
        filter_pixel(job->src, job->dst, x, y);
  }
  return NULL;
}

/* denoiser.c:129-149 */
void oracle_denoise_image(Image const *src, Image const *dst, isize n_threads) {
  assert(src->pixels.data != dst->pixels.data);
  assert(src->width == dst->width && src->height == dst->height);
  if (n_threads < 1) n_threads = 1;
  Job job;
  job.src = src;
  job.dst = dst;
  atomic_init(&job.next_chunk, 0);
  pthread_t *th = malloc(sizeof(pthread_t) * (size_t)n_threads);
  for (isize i = 1; i < n_threads; i++) pthread_create(&th[i], NULL, worker, &job);
  worker(&job);
  for (isize i = 1; i < n_threads; i++) pthread_join(th[i], NULL);
  free(th);
}
