/* driver.c — the C host program, a drop-in for the reference's driver
 * (reference driver.c:420-508 flags, :730-878 flow): same flags
 *   -W <width> -H <height> -S <samples> -T <threads> -B <max_bounces> -V -D -O <out.(png|qoi|ppm)> <model>
 * same defaults (1024x1024, 16 spp, 8 bounces, 1 thread, output.png), same
 * progress bar and verbose prints, same call sequence into scene_init /
 * render_thread_proc / rendering_context_is_finished / denoise_image — which
 * here resolve to the host BVH builder and the sm_100a library.
 *
 * Additions (long options, none collide with the reference's flags):
 *   --env <image>      environment map (default ./background.png; if that file
 *                      is absent the procedural stand-in is used — the reference
 *                      exits instead, but its background.png is not in the tree)
 *   --eye x,y,z --target x,y,z [--fov degrees]   camera override
 *   --seed <u32>       per-(pixel,sample) seed
 *   --device <n>       CUDA device
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

#include "rt_gpu.h"
#include "rt_host.h"

typedef struct {
  char const *model, *output_path, *env_path;
  isize width, height, samples, max_bounces, n_threads;
  bool verbose, denoise, has_eye, has_target;
  f32 eye[3], target[3], fov_degrees;
  u32 seed;
  int device;
} Config;

static void print_usage(char const *argv0) {
  fprintf(stderr, "%s -W <width> -H <height> -S <samples> -T <threads> -B <max_bounces> <model.(obj|glb|gltf)> -O output.(qoi|png|ppm)\n", argv0);
}

static f64 now_ms(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

static bool parse_vec3(char const *s, f32 out[3]) { return sscanf(s, "%f,%f,%f", &out[0], &out[1], &out[2]) == 3; }

/* driver.c:444-508 */
static bool parse_args(int argc, char **argv, Config *c) {
  for (int i = 1; i < argc;) {
    char const *arg = argv[i];
    if (!strncmp(arg, "--", 2)) {
      if (i == argc - 1) { print_usage(argv[0]); return false; }
      char const *val = argv[i + 1];
      if      (!strcmp(arg, "--env"))    c->env_path = val;
      else if (!strcmp(arg, "--eye"))    c->has_eye = parse_vec3(val, c->eye);
      else if (!strcmp(arg, "--target")) c->has_target = parse_vec3(val, c->target);
      else if (!strcmp(arg, "--fov"))    c->fov_degrees = (f32)atof(val);
      else if (!strcmp(arg, "--seed"))   c->seed = (u32)strtoul(val, NULL, 10);
      else if (!strcmp(arg, "--device")) c->device = atoi(val);
      else { print_usage(argv[0]); return false; }
      i += 2;
      continue;
    }
    if (arg[0] == '-') {
      if (strlen(arg) != 2) { print_usage(argv[0]); return false; }
      if (arg[1] == 'V') { c->verbose = true; i += 1; continue; }
      if (arg[1] == 'D') { c->denoise = true; i += 1; continue; }
      if (i == argc - 1) { print_usage(argv[0]); return false; }
      if (arg[1] == 'O') { c->output_path = argv[i + 1]; i += 2; continue; }
      isize number = atol(argv[i + 1]);
      switch (arg[1]) {
        case 'W': c->width = number; break;
        case 'H': c->height = number; break;
        case 'S': c->samples = number; break;
        case 'T': c->n_threads = number; break;
        case 'B': c->max_bounces = number; break;
        default: print_usage(argv[0]); return false;
      }
      i += 2;
    } else {
      if (c->model) { print_usage(argv[0]); return false; }
      c->model = arg;
      i += 1;
    }
  }
  if (!c->model) { print_usage(argv[0]); return false; }
  return true;
}

static void *render_entry(void *ctx) { render_thread_proc((Rendering_Context *)ctx); return NULL; }

int main(int argc, char **argv) {
  f64 t_process = now_ms();
  /* driver.c:733-742 */
  Config config = { .width = 1024, .height = 1024, .samples = 16, .max_bounces = 8, .n_threads = 1,
                    .output_path = "output.png", .env_path = "background.png", .fov_degrees = 70.0f };
  if (!parse_args(argc, argv, &config)) return 1;
  if (rt_gpu_init(config.device)) { fprintf(stderr, "%s\n", rt_gpu_last_error()); return 1; }

  Image image = rt_image_alloc(config.width, config.height, 3);

  Scene scene;
  memset(&scene, 0, sizeof scene);
  Image background;
  if (access(config.env_path, R_OK) == 0) {
    if (!rt_load_texture(config.env_path, &background)) { fprintf(stderr, "Failed to load texture: '%s'\n", config.env_path); return 1; }
  } else {
    fprintf(stderr, "'%s' not found: using the procedural environment\n", config.env_path);
    rt_generate_background(&background, 2048, 1024);
  }
  scene.background.data = &background;
  scene.background.proc = rt_gpu_background_proc;
  rt_gpu_register_background(rt_gpu_background_proc);
  rt_gpu_register_pbr_shader(rt_gpu_pbr_shader_proc);

  rt_camera_default(&scene.camera);
  RT_Model model;
  if (!rt_load_model_file(config.model, rt_gpu_pbr_shader_proc, &model, &scene.camera)) return 1;
  if (config.has_eye && config.has_target) {
    Vec3 eye = {{ config.eye[0], config.eye[1], config.eye[2] }}, target = {{ config.target[0], config.target[1], config.target[2] }}, up = {{ 0, 1, 0 }};
    rt_camera_look_at(&scene.camera, eye, target, up, config.fov_degrees * 3.14159265f / 180.0f);
  }

  f64 t_bvh = now_ms();
  scene_init(&scene, model.triangles);
  if (config.verbose) {
    printf("Bvh generated in %ldms\n", (long)(now_ms() - t_bvh));
    printf("Width:     %ld\nHeight:    %ld\nSamples:   %ld\nBounces:   %ld\nThreads:   %ld\n",
           (long)config.width, (long)config.height, (long)config.samples, (long)config.max_bounces, (long)config.n_threads);
    printf("BVH-Nodes: %ld\nBVH-Depth: %ld\nTriangles: %ld\n\n", (long)scene.bvh.nodes.len, (long)scene.bvh.depth, (long)model.triangles.len);
  }
  if (rt_gpu_scene_upload(&scene)) { fprintf(stderr, "%s\n", rt_gpu_last_error()); return 1; }
  RT_GPU_Options options;
  rt_gpu_get_options(&options);
  options.user_seed = config.seed;
  rt_gpu_set_options(&options);

  f64 t_render = now_ms();
  Rendering_Context ctx;
  memset(&ctx, 0, sizeof ctx);
  ctx.image = image;
  ctx.scene = &scene;
  ctx.max_bounces = config.max_bounces;
  ctx.samples = config.samples;
  ctx.n_threads = (i32)config.n_threads;

  pthread_t *threads = malloc(sizeof(pthread_t) * (size_t)config.n_threads);
  for (isize i = 0; i < config.n_threads; i++) pthread_create(&threads[i], NULL, render_entry, &ctx);

  isize n_chunks = ((image.width + RT_CHUNK_SIZE - 1) / RT_CHUNK_SIZE) * ((image.height + RT_CHUNK_SIZE - 1) / RT_CHUNK_SIZE);
  char const *bar = "====================";
  while (!rendering_context_is_finished(&ctx)) {
    f32 p = (f32)ctx._current_chunk / (f32)n_chunks;
    if (p > 1.0f) p = 1.0f;
    printf("\r[%-20.*s] %d%%", (int)(p * 20), bar, (int)(100 * p));
    fflush(stdout);
    usleep(20 * 1000);          /* the reference sleeps 500 ms (driver.c:817), which would quantise GPU timings */
  }
  printf("\r[%s] 100%%\n", bar);
  for (isize i = 0; i < config.n_threads; i++) pthread_join(threads[i], NULL);
  free(threads);

  f64 render_ms = now_ms() - t_render;
  printf("%ldms\n", (long)render_ms);
  if (config.verbose) {
    printf("%ld samples/second\n", (long)((f64)config.width * (f64)config.height * (f64)config.samples / (render_ms / 1e3)));
    printf("GPU trace kernels: %.3fms in %d launches\n", rt_gpu_last_kernel_ms(), rt_gpu_last_launches());
  }

  if (config.denoise) {
    f64 t0 = now_ms();
    Image denoised = rt_image_alloc(image.width, image.height, image.components);
    denoise_image(&image, &denoised, config.n_threads);
    rt_image_free(&image);
    image = denoised;
    printf("Denoising: %ldms\n", (long)(now_ms() - t0));
  }

  f64 t_out = now_ms();
  if (!rt_save_image(config.output_path, &image)) { fprintf(stderr, "Failed to encode output file\n"); return 1; }
  if (config.verbose) {
    printf("Output file written in %ldms\n", (long)(now_ms() - t_out));
    printf("Time to image: %ldms\n", (long)(now_ms() - t_process));
  }
  return 0;
}
