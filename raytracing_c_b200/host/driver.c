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
 *   --device <n>       CUDA device (first of --gpus)
 *   --gpus <n>         devices this process drives (rt_gpu_init_devices): the frame is split over them, the
 *                      accumulators are combined on the first one before the resolve and the denoise pass
 *   --split auto|samples|chunks   how (RT_GPU_Options.split_mode)      --reduce p2p|nccl   with what
 *   --dump-accum <file>    raw f32 W*H*3 sums of cast_ray (pre-division), little endian
 *   --dump-hit-ids <file>  raw i32 W*H primary-hit slots of sample 0 (-1 = miss)
 *   --scene-cache <file>   load the BVH and triangle block from this file if it exists (and matches the model's
 *                      triangle count), else build them and write it (reference scene.c:18-76; host/scene_cache.c)
 *   --gpu-decode 1     JPEG textures stay compressed on the host and are decoded by nvJPEG during the scene upload
 *                      (not byte-identical to the host decoder: INTEGRATION.md)
 *   --pinned 1         host buffers (texels, nodes, triangles, image) in pinned memory, DMA-read in place by every
 *                      upload (worth it for a host that re-uploads per frame; pinning ~60 MB costs more than
 *                      one staged upload, so a one-shot render keeps them pageable)
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
  int device, gpus, split_mode, reduce_mode, pinned, gpu_decode;
  char const *dump_accum, *dump_hit_ids, *scene_cache;
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
      else if (!strcmp(arg, "--gpus"))   c->gpus = atoi(val);
      else if (!strcmp(arg, "--pinned")) c->pinned = atoi(val);
      else if (!strcmp(arg, "--gpu-decode")) c->gpu_decode = atoi(val);
      else if (!strcmp(arg, "--scene-cache"))  c->scene_cache = val;
      else if (!strcmp(arg, "--dump-accum"))   c->dump_accum = val;
      else if (!strcmp(arg, "--dump-hit-ids")) c->dump_hit_ids = val;
      else if (!strcmp(arg, "--split"))
        c->split_mode = !strcmp(val, "samples") ? RT_GPU_SPLIT_SAMPLES : !strcmp(val, "chunks") ? RT_GPU_SPLIT_CHUNKS : RT_GPU_SPLIT_AUTO;
      else if (!strcmp(arg, "--reduce")) c->reduce_mode = !strcmp(val, "nccl") ? RT_GPU_REDUCE_NCCL : RT_GPU_REDUCE_P2P;
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

typedef struct { int n; int *devices; int status; int joined; f64 done_ms; isize width, height, samples, bounces; } Init_Job;
static void *init_entry(void *p) {
  Init_Job *job = p;
  job->status = rt_gpu_init_devices(job->n, job->devices);
  if (!job->status) rt_gpu_prepare_frame(job->width, job->height, job->samples, job->bounces);      /* best effort */
  job->done_ms = now_ms();
  return NULL;
}

static void *render_entry(void *ctx) { render_thread_proc((Rendering_Context *)ctx); return NULL; }

int main(int argc, char **argv) {
  f64 t_process = now_ms();
  /* driver.c:733-742 */
  Config config = { .width = 1024, .height = 1024, .samples = 16, .max_bounces = 8, .n_threads = 1,
                    .output_path = "output.png", .env_path = "background.png", .fov_degrees = 70.0f, .gpus = 1 };
  if (!parse_args(argc, argv, &config)) return 1;
  if (config.gpus < 1 || config.gpus > 16) { fprintf(stderr, "--gpus must be between 1 and 16\n"); return 1; }
  int devices[16];
  for (int i = 0; i < config.gpus; i++) devices[i] = config.device + i;
  /* Creating the CUDA context dominates time-to-image (1-2 s of a 2.2 s run on an 8-GPU box).  The driver initialises
   * every VISIBLE device when the first context is made, so unless the user already chose, show it only the ones used. */
  if (!getenv("CUDA_VISIBLE_DEVICES")) {
    char list[128] = "";
    for (int i = 0; i < config.gpus; i++) snprintf(list + strlen(list), sizeof list - strlen(list), i ? ",%d" : "%d", devices[i]);
    setenv("CUDA_VISIBLE_DEVICES", list, 1);
    for (int i = 0; i < config.gpus; i++) devices[i] = i;
  }
  /* CUDA context creation (1-2 s) runs on its own thread while this one loads the model, decodes the textures and
   * builds the BVH; pinned host buffers need the context, so --pinned 1 waits for it first */
  /* options first: they need no device, and the init thread sizes its allocations by them */
  RT_GPU_Options options;
  rt_gpu_get_options(&options);
  options.user_seed = config.seed;
  options.split_mode = config.split_mode;
  options.reduce_mode = config.reduce_mode;
  options.keep_hit_ids = config.dump_hit_ids != NULL;
  rt_gpu_set_options(&options);
  Init_Job init = { config.gpus, devices, 0, 0, 0, config.width, config.height, config.samples, config.max_bounces };
  pthread_t init_thread;
  pthread_create(&init_thread, NULL, init_entry, &init);
  if (config.pinned) {
    pthread_join(init_thread, NULL);
    init.joined = 1;
    if (init.status) { fprintf(stderr, "%s\n", rt_gpu_last_error()); return 1; }
    rt_host_set_buffer_allocator(rt_gpu_host_alloc, rt_gpu_host_free);
  }
  if (config.gpu_decode) rt_host_defer_jpeg_decode(true);

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

  rt_camera_default(&scene.camera);
  RT_Model model;
  f64 t_load = now_ms();
  if (!rt_load_model_file(config.model, rt_gpu_pbr_shader_proc, &model, &scene.camera)) return 1;
  f64 t_load_done = now_ms();
  if (config.has_eye && config.has_target) {
    Vec3 eye = {{ config.eye[0], config.eye[1], config.eye[2] }}, target = {{ config.target[0], config.target[1], config.target[2] }}, up = {{ 0, 1, 0 }};
    rt_camera_look_at(&scene.camera, eye, target, up, config.fov_degrees * 3.14159265f / 180.0f);
  }

  f64 t_bvh = now_ms();
  bool cached = false;
  if (config.scene_cache && access(config.scene_cache, R_OK) == 0) {
    Camera keep = scene.camera;       /* the command line and the model decide the camera, not the cache */
    cached = scene_load_file(config.scene_cache, &scene, model.materials, model.n_materials, rt_gpu_pbr_shader_proc);
    scene.camera = keep;
    if (!cached) fprintf(stderr, "scene cache '%s' not used: %s\n", config.scene_cache, rt_host_last_error());
  }
  if (!cached) {
    scene_init(&scene, model.triangles);
    if (config.scene_cache && !scene_save_file(config.scene_cache, &scene, model.materials, model.n_materials))
      fprintf(stderr, "scene cache '%s' not written: %s\n", config.scene_cache, rt_host_last_error());
  }
  if (config.verbose) {
    printf("Bvh %s in %ldms\n", cached ? "loaded from the scene cache" : "generated", (long)(now_ms() - t_bvh));
    printf("Width:     %ld\nHeight:    %ld\nSamples:   %ld\nBounces:   %ld\nThreads:   %ld\n",
           (long)config.width, (long)config.height, (long)config.samples, (long)config.max_bounces, (long)config.n_threads);
    printf("BVH-Nodes: %ld\nBVH-Depth: %ld\nTriangles: %ld\n\n", (long)scene.bvh.nodes.len, (long)scene.bvh.depth, (long)model.triangles.len);
  }
  if (!init.joined) {
    pthread_join(init_thread, NULL);
    if (init.status) { fprintf(stderr, "%s\n", rt_gpu_last_error()); return 1; }
  }
  f64 t_init_done = init.done_ms;
  rt_gpu_register_background(rt_gpu_background_proc);      /* (these take the library's lock: after the init thread) */
  rt_gpu_register_pbr_shader(rt_gpu_pbr_shader_proc);
  f64 t_upload = now_ms();
  if (rt_gpu_scene_upload(&scene)) { fprintf(stderr, "%s\n", rt_gpu_last_error()); return 1; }
  if (config.verbose)
    printf("GPU init %ldms (on its own thread, beside the model load), model load + texture decode %ldms, scene upload issued in %ldms\n", (long)(t_init_done - t_process),
           (long)(t_load_done - t_load), (long)(now_ms() - t_upload));

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
  if (rt_gpu_last_status()) { fprintf(stderr, "render failed: %s\n", rt_gpu_last_error()); return 1; }
  printf("%ldms\n", (long)render_ms);
  if (config.verbose) {
    printf("%ld samples/second\n", (long)((f64)config.width * (f64)config.height * (f64)config.samples / (render_ms / 1e3)));
    f64 parts[4];
    rt_gpu_last_frame_breakdown(parts);
    printf("GPUs: %d (%s split), render kernels %.3fms in %d launches, reduce+resolve %.3fms, image D2H %.3fms\n", rt_gpu_device_count(),
           parts[3] == RT_GPU_SPLIT_CHUNKS ? "chunk" : "sample", rt_gpu_last_kernel_ms(), rt_gpu_last_launches(), parts[1], parts[2]);
  }
  if (config.dump_accum || config.dump_hit_ids) {
    size_t n = (size_t)config.width * (size_t)config.height;
    if (config.dump_accum) {
      f32 *accum = malloc(n * 3 * sizeof(f32));
      FILE *f = fopen(config.dump_accum, "wb");
      if (!accum || !f || rt_gpu_read_accum(accum, (isize)(n * 3)) || fwrite(accum, sizeof(f32), n * 3, f) != n * 3) {
        fprintf(stderr, "--dump-accum failed: %s\n", rt_gpu_last_error());
        return 1;
      }
      fclose(f);
      free(accum);
    }
    if (config.dump_hit_ids) {
      i32 *ids = malloc(n * sizeof(i32));
      FILE *f = fopen(config.dump_hit_ids, "wb");
      if (!ids || !f || rt_gpu_read_hit_ids(ids, (isize)n) || fwrite(ids, sizeof(i32), n, f) != n) {
        fprintf(stderr, "--dump-hit-ids failed: %s\n", rt_gpu_last_error());
        return 1;
      }
      fclose(f);
      free(ids);
    }
  }

  if (config.denoise) {
    f64 t0 = now_ms();
    Image denoised = rt_image_alloc(image.width, image.height, image.components);
    denoise_image(&image, &denoised, config.n_threads);
    if (rt_gpu_last_status()) { fprintf(stderr, "denoise failed: %s\n", rt_gpu_last_error()); return 1; }
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
