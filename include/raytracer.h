/* raytracer.h — the path-tracing entry points, kept exactly as the reference
 * declares them (reference raytracer.h:44-56) so a C host written against the
 * reference links against libraytracer_gpu.so unchanged.
 *
 * Threading contract (reference driver.c:793-818, raytracer.c:619-627,786-794):
 * the host starts `n_threads` OS threads that each call render_thread_proc with
 * the SAME context, then polls rendering_context_is_finished and reads
 * _current_chunk for its progress bar.  In this implementation the thread that
 * claims chunk 0 owns the GPU launch; it advances _current_chunk as sample
 * slices complete and is the last one to decrement n_threads, so n_threads==0
 * still means "image.pixels is complete".
 */
#ifndef RT_RAYTRACER_H
#define RT_RAYTRACER_H

#include "scene.h"

#ifdef __cplusplus
extern "C" {
#define RT_ATOMIC_I32 volatile i32
#else
#include <stdatomic.h>
#define RT_ATOMIC_I32 _Atomic i32
#endif

typedef struct {
  Vec3 position;
  Vec3 direction;
} Ray;

typedef struct {
  Image         image;
  Scene        *scene;
  isize         samples, max_bounces;
  RT_ATOMIC_I32 n_threads, _current_chunk;
} Rendering_Context;

/* reference raytracer.h:51-56 */
extern void rendering_context_finish(Rendering_Context *context);
extern bool rendering_context_is_finished(Rendering_Context *context);
extern void render_thread_proc(Rendering_Context *context);
extern void lightmap_bake(Image const *lightmap, Scene const *scene, isize samples);

#ifdef __cplusplus
}
#endif
#endif
