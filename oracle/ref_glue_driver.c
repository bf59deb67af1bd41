/* ref_glue_driver.c — TEST INFRASTRUCTURE.  Includes the reference's driver.c
 * UNMODIFIED (main is renamed on the command line and never called) and exports
 * its file-local shading callbacks and its translation unit's random_state
 * (common.h:13 is `static thread_local` in a header, so driver.c has its own). */
#include "driver.c"

Shader_Proc     ref_shader_proc(void)     { return disney_shader_proc; }
Background_Proc ref_background_proc(void) { return (Background_Proc)sample_background; }
u32            *ref_random_state(void)    { return &random_state; }
Color3 ref_sample_texture_bilinear(Image const *texture, Vec2 uv) { return sample_texture_bilinear(texture, uv); }
Color3 ref_sample_background(Image const *image, Vec3 dir)        { return sample_background(image, dir); }
