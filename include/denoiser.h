/* denoiser.h — reference denoiser.h:4.  Synchronous; `n_threads` is advisory
 * on the GPU (the stencil kernel covers the image in one launch). */
#ifndef RT_DENOISER_H
#define RT_DENOISER_H

#include "rt_base.h"

#ifdef __cplusplus
extern "C" {
#endif

extern void denoise_image(Image const *src, Image const *dst, isize n_threads);

#ifdef __cplusplus
}
#endif
#endif
