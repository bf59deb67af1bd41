/* rt_base.h — Codin-free base types for the raytracer.h / denoiser.h boundary.
 *
 * The reference pulls every scalar, vector, slice and image type from the
 * author's "Codin" stdlib (reference common.h:3-4, scene.h:3-6), which is not
 * vendored.  The field layout of those types is therefore unpinned; this header
 * defines the layout used on both sides of the C-ABI in this repo.  Only the
 * fields the hot path touches are present (reference driver.c:747-754,
 * denoiser.c:16-38, raytracer.c:612,670-672).
 */
#ifndef RT_BASE_H
#define RT_BASE_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef float     f32;
typedef double    f64;
typedef int32_t   i32;
typedef int64_t   i64;
typedef uint32_t  u32;
typedef uint64_t  u64;
typedef uint8_t   u8;
typedef uint16_t  u16;
typedef ptrdiff_t isize;
typedef void     *rawptr;

typedef union { struct { f32 x, y; };          f32 data[2]; } Vec2;
typedef union { struct { f32 x, y, z; };       struct { f32 r, g, b; }; f32 data[3]; } Vec3;
typedef union { struct { f32 x, y, z, w; };    struct { f32 r, g, b, a; }; f32 data[4]; } Vec4;
typedef Vec3 Color3;
typedef Vec4 Color4;

/* Row-major, rows[r][c]; a camera's view_matrix is camera-to-world
 * (reference raytracer.c:612,670-672). */
typedef struct { f32 rows[4][4]; } Matrix_4x4;

typedef struct { u8 *data; isize len; } Byte_Slice;

/* PT_RT_JPEG_BYTES is this repo's extension: pixels holds a COMPRESSED baseline JPEG (pixels.len bytes), width and
 * height come from its frame header, components is 3.  Only the GPU library's scene upload accepts such an image
 * (it decodes with nvJPEG straight into device memory, rt_gpu.h); librt_host produces it on request. */
enum { PT_u8 = 0, PT_RT_JPEG_BYTES = 0x4A50 };

/* stride is in pixels (reference raytracer.c:714, denoiser.c:24). */
typedef struct {
  Byte_Slice pixels;
  isize      width, height, stride;
  i32        components;
  i32        pixel_type;
} Image;

#define RT_EPSILON 0.0001f   /* reference common.h:8 */
#define RT_SIMD_WIDTH 8      /* reference raytracer.h:6 */
#define RT_CHUNK_SIZE 32     /* reference raytracer.c:601, denoiser.c:45 */

#ifdef __cplusplus
}
#endif
#endif
