/* ref_glue_common.c — TEST INFRASTRUCTURE.  Runtime pieces of the Codin stand-in
 * (oracle/codin_shim) and the entry points tests use to drive the UNMODIFIED
 * reference sources compiled into oracle/_ref/libref.so.  Loaders, encoders and
 * file IO are stubs: the reference's main() and model loading are never run —
 * triangles come from this repo's host loaders, exactly as for the oracle. */
#include "codin/codin.h"
/* the reference's raytracer.h (found via -I$(REF)); it defines SIMD_WIDTH before including scene.h */
#include "raytracer.h"

thread_local Codin_Context context;
String_Slice os_args;

typedef struct { Thread_Proc proc; rawptr arg; } Thread_Start;
static void *thread_trampoline(void *p) {
  Thread_Start s = *(Thread_Start *)p;
  free(p);
  s.proc(s.arg);
  return NULL;
}
void thread_create(Thread_Proc proc, rawptr arg, isize stack, isize tls) {
  (void)stack; (void)tls;
  Thread_Start *s = malloc(sizeof *s);
  s->proc = proc; s->arg = arg;
  pthread_t t;
  pthread_create(&t, NULL, thread_trampoline, s);
  pthread_detach(t);
}

Fd_Result    file_open(String path, int flags) { (void)path; (void)flags; Fd_Result r = { -1, 1 }; return r; }
Writer       writer_from_handle(Fd fd) { Writer w = { fd }; return w; }
Bytes_Result read_entire_file_path(String path, Allocator a) { (void)path; (void)a; Bytes_Result r = { { 0, 0 }, 1 }; return r; }
void         write_bytes(Writer const *w, Byte_Slice bytes) { (void)w; (void)bytes; }
bool png_save_writer(Writer const *w, Image const *image) { (void)w; (void)image; return false; }
bool qoi_save_writer(Writer const *w, Image const *image) { (void)w; (void)image; return false; }
bool ppm_save_writer(Writer const *w, Image const *image) { (void)w; (void)image; return false; }
bool stb_image_load_bytes(Byte_Slice bytes, Image *image, Allocator a) { (void)bytes; (void)image; (void)a; return false; }
bool obj_load(String text, Obj_File *obj, bool flag, Allocator a) { (void)text; (void)obj; (void)flag; (void)a; return false; }
bool gltf_parse(Byte_Slice data, String path, Gltf_File *gltf, Allocator a) { (void)data; (void)path; (void)gltf; (void)a; return false; }
bool gltf_load_buffers(String path, Gltf_File *gltf, Allocator a) { (void)path; (void)gltf; (void)a; return false; }
void gltf_to_triangles(Gltf_File *gltf, Gltf_Triangle_Vector *out) { (void)gltf; (void)out; }

/* reference scene.c:416 — sorts the caller's slice in place, so work on a copy */
void ref_scene_init(Scene *scene, Triangle_Slice src) {
  Triangle_Slice copy;
  copy.len = src.len;
  copy.data = malloc(sizeof(Triangle) * (usize)(src.len ? src.len : 1));
  memcpy(copy.data, src.data, sizeof(Triangle) * (usize)src.len);
  scene_init(scene, copy, context.allocator);
  free(copy.data);
}

isize ref_sizeof_triangle(void)     { return size_of(Triangle); }
isize ref_sizeof_triangle_aos(void) { return size_of(Triangle_AOS); }
isize ref_sizeof_scene(void)        { return size_of(Scene); }
isize ref_sizeof_context(void)      { return size_of(Rendering_Context); }
isize ref_sizeof_bvh_node(void)     { return size_of(BVH_Node); }
