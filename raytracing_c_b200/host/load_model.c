/* load_model.c — OBJ/MTL and glTF 2.0 (.gltf / .glb) ingestion into
 * Triangle_Slice + PBR_Shader_Data[] + Image[] (+ camera), the inputs of
 * scene_init.  Mirrors what the reference's driver.c builds through Codin's
 * obj_load / gltf_parse / gltf_to_triangles (driver.c:510-728).
 *
 * UNPINNED (Codin not in tree): triangle emission order (here: file order for
 * OBJ; node-index order, then primitive, then index order for glTF), OBJ V
 * flip (here v' = 1 - v, the usual convention for image-space sampling),
 * normal transform (here inverse-transpose of the node's linear part).
 * glTF defaults follow the glTF 2.0 specification.
 */
#include <ctype.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "rt_host.h"
#include "rt_json.h"

#include <pthread.h>

/* one texture decode, run on its own thread (see the images loop of the glTF loader) */
typedef struct {
  u8 const *bytes; size_t len;      /* embedded image, or */
  char     *path;                   /* file next to the model */
  Image    *out;
  bool      ok, threaded;
  pthread_t thread;
  char      error[256];
} Decode_Job;

static void *decode_job_run(void *arg) {
  Decode_Job *job = arg;
  if (job->error[0]) return NULL;
  job->ok = job->bytes ? rt_image_decode(job->bytes, job->len, job->out) : rt_load_texture(job->path, job->out);
  if (!job->ok) snprintf(job->error, sizeof job->error, "%s", rt_host_last_error());   /* thread-local on the worker */
  return NULL;
}


void rt_host_set_error(char const *msg);

static u8 *slurp(char const *path, size_t *len) {
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

static char *sibling_path(char const *base, char const *name) {
  char const *slash = strrchr(base, '/');
  size_t dir = slash ? (size_t)(slash - base + 1) : 0;
  char *out = malloc(dir + strlen(name) + 1);
  memcpy(out, base, dir);
  strcpy(out + dir, name);
  return out;
}

/* driver.c:550-566 non-PBR default: base = Kd, roughness 0.5 */
static PBR_Shader_Data default_material(void) {
  PBR_Shader_Data m;
  memset(&m, 0, sizeof m);
  m.base_color.x = m.base_color.y = m.base_color.z = 0.8f;
  m.roughness = 0.5f;
  return m;
}

/* ===================================================================== OBJ */
typedef struct {
  char  name[128];
  f32   kd[3], ke[3];
  bool  is_pbr;
  f32   pr, pm, ps, aniso;
  char  map_kd[256], map_ke[256], map_pr[256], map_pm[256], map_norm[256];
} Mtl;

typedef struct { Mtl *items; isize n, cap; } Mtl_List;

static char *next_line(char **cursor) {
  char *s = *cursor;
  if (!*s) return NULL;
  char *e = s;
  while (*e && *e != '\n') e++;
  if (*e) { *e = 0; *cursor = e + 1; } else *cursor = e;
  size_t n = strlen(s);
  while (n && (s[n - 1] == '\r' || s[n - 1] == ' ' || s[n - 1] == '\t')) s[--n] = 0;
  while (*s == ' ' || *s == '\t') s++;
  return s;
}

static void copy_token(char *dst, size_t cap, char const *src) {
  while (*src == ' ' || *src == '\t') src++;
  /* texture statements may carry options (-bm 1.0 ...); take the last token */
  char const *last = strrchr(src, ' ');
  if (last && src[0] == '-') src = last + 1;
  snprintf(dst, cap, "%s", src);
}

static void parse_mtl(char const *path, Mtl_List *list) {
  size_t len;
  char *text = (char *)slurp(path, &len);
  if (!text) return;
  char *cursor = text, *line;
  Mtl *cur = NULL;
  while ((line = next_line(&cursor))) {
    if (!strncmp(line, "newmtl ", 7)) {
      if (list->n == list->cap) { list->cap = list->cap ? list->cap * 2 : 8; list->items = realloc(list->items, sizeof(Mtl) * (size_t)list->cap); }
      cur = &list->items[list->n++];
      memset(cur, 0, sizeof *cur);
      snprintf(cur->name, sizeof cur->name, "%s", line + 7);
      cur->kd[0] = cur->kd[1] = cur->kd[2] = 0.8f;
      cur->pr = 0.5f;
      continue;
    }
    if (!cur) continue;
    if      (!strncmp(line, "Kd ", 3)) sscanf(line + 3, "%f %f %f", &cur->kd[0], &cur->kd[1], &cur->kd[2]);
    else if (!strncmp(line, "Ke ", 3)) sscanf(line + 3, "%f %f %f", &cur->ke[0], &cur->ke[1], &cur->ke[2]);
    else if (!strncmp(line, "Pr ", 3)) { cur->is_pbr = true; sscanf(line + 3, "%f", &cur->pr); }
    else if (!strncmp(line, "Pm ", 3)) { cur->is_pbr = true; sscanf(line + 3, "%f", &cur->pm); }
    else if (!strncmp(line, "Ps ", 3)) { cur->is_pbr = true; sscanf(line + 3, "%f", &cur->ps); }
    else if (!strncmp(line, "aniso ", 6)) { cur->is_pbr = true; sscanf(line + 6, "%f", &cur->aniso); }
    else if (!strncmp(line, "map_Kd ", 7)) copy_token(cur->map_kd, sizeof cur->map_kd, line + 7);
    else if (!strncmp(line, "map_Ke ", 7)) copy_token(cur->map_ke, sizeof cur->map_ke, line + 7);
    else if (!strncmp(line, "map_Pr ", 7)) { cur->is_pbr = true; copy_token(cur->map_pr, sizeof cur->map_pr, line + 7); }
    else if (!strncmp(line, "map_Pm ", 7)) { cur->is_pbr = true; copy_token(cur->map_pm, sizeof cur->map_pm, line + 7); }
    else if (!strncmp(line, "norm ", 5))   { cur->is_pbr = true; copy_token(cur->map_norm, sizeof cur->map_norm, line + 5); }
  }
  free(text);
}

typedef struct { char path[256]; isize image; } Tex_Ref;

static Image *texture_for(RT_Model *model, Tex_Ref **refs, isize *n_refs, char const *obj_path, char const *name) {
  if (!name[0]) return NULL;
  for (isize i = 0; i < *n_refs; i++) if (!strcmp((*refs)[i].path, name)) return &model->images[(*refs)[i].image];
  char *full = sibling_path(obj_path, name);
  Image im;
  bool ok = rt_load_texture(full, &im);
  if (!ok) { fprintf(stderr, "Failed to load texture: '%s'\n", full); free(full); return NULL; }
  free(full);
  /* images[] was sized for the worst case up front so pointers stay valid */
  model->images[model->n_images] = im;
  *refs = realloc(*refs, sizeof(Tex_Ref) * (size_t)(*n_refs + 1));
  snprintf((*refs)[*n_refs].path, sizeof (*refs)[*n_refs].path, "%s", name);
  (*refs)[*n_refs].image = model->n_images;
  (*n_refs)++;
  return &model->images[model->n_images++];
}

static bool parse_face_vertex(char const **cursor, int *v, int *vt, int *vn) {
  char const *s = *cursor;
  while (*s == ' ' || *s == '\t') s++;
  if (!*s) return false;
  char *end;
  *v = (int)strtol(s, &end, 10); *vt = 0; *vn = 0;
  if (end == s) return false;
  s = end;
  if (*s == '/') {
    s++;
    if (*s != '/') { *vt = (int)strtol(s, &end, 10); s = end; }
    if (*s == '/') { s++; *vn = (int)strtol(s, &end, 10); s = end; }
  }
  *cursor = s;
  return true;
}

static bool load_obj(char const *path, u8 *data, Shader_Proc proc, RT_Model *model) {
  typedef struct { f32 *p; isize n, cap; } Floats;
  Floats pos = {0}, tex = {0}, nrm = {0};
  #define PUSH3(arr, a, b, c) do { if (arr.n + 3 > arr.cap) { arr.cap = arr.cap ? arr.cap * 2 : 1024; arr.p = realloc(arr.p, sizeof(f32) * (size_t)arr.cap); } \
                                   arr.p[arr.n++] = a; arr.p[arr.n++] = b; arr.p[arr.n++] = c; } while (0)
  Mtl_List mtls = {0};
  Triangle *tris = NULL; isize n_tris = 0, cap_tris = 0;
  isize *tri_mtl = NULL;
  isize cur_mtl = -1;
  bool missing_mtl_warned = false;

  char *cursor = (char *)data, *line;
  while ((line = next_line(&cursor))) {
    if (line[0] == 'v' && line[1] == ' ') { f32 a = 0, b = 0, c = 0; sscanf(line + 2, "%f %f %f", &a, &b, &c); PUSH3(pos, a, b, c); }
    else if (line[0] == 'v' && line[1] == 't') { f32 a = 0, b = 0; sscanf(line + 3, "%f %f", &a, &b); PUSH3(tex, a, 1.0f - b, 0); }
    else if (line[0] == 'v' && line[1] == 'n') { f32 a = 0, b = 0, c = 0; sscanf(line + 3, "%f %f %f", &a, &b, &c); PUSH3(nrm, a, b, c); }
    else if (!strncmp(line, "mtllib ", 7)) { char *p = sibling_path(path, line + 7); parse_mtl(p, &mtls); free(p); }
    else if (!strncmp(line, "usemtl ", 7)) {
      cur_mtl = -1;
      for (isize i = 0; i < mtls.n; i++) if (!strcmp(mtls.items[i].name, line + 7)) cur_mtl = i;
      if (cur_mtl < 0 && !missing_mtl_warned) { fprintf(stderr, "material '%s' not found, using the default material\n", line + 7); missing_mtl_warned = true; }
    }
    else if (line[0] == 'f' && line[1] == ' ') {
      int v[64], vt[64], vn[64], n = 0;
      char const *c = line + 2;
      while (n < 64 && parse_face_vertex(&c, &v[n], &vt[n], &vn[n])) n++;
      for (int k = 1; k + 1 < n; k++) {
        if (n_tris == cap_tris) { cap_tris = cap_tris ? cap_tris * 2 : 4096; tris = realloc(tris, sizeof(Triangle) * (size_t)cap_tris); tri_mtl = realloc(tri_mtl, sizeof(isize) * (size_t)cap_tris); }
        Triangle *t = &tris[n_tris];
        memset(t, 0, sizeof *t);
        int corner[3] = { 0, k, k + 1 };
        for (int j = 0; j < 3; j++) {
          int iv = v[corner[j]], it = vt[corner[j]], in = vn[corner[j]];
          if (iv < 0) iv = (int)(pos.n / 3) + iv + 1;
          if (it < 0) it = (int)(tex.n / 3) + it + 1;
          if (in < 0) in = (int)(nrm.n / 3) + in + 1;
          if (iv >= 1 && iv <= pos.n / 3) { t->positions[j].x = pos.p[3 * (iv - 1)]; t->positions[j].y = pos.p[3 * (iv - 1) + 1]; t->positions[j].z = pos.p[3 * (iv - 1) + 2]; }
          if (it >= 1 && it <= tex.n / 3) { t->tex_coords[j].x = tex.p[3 * (it - 1)]; t->tex_coords[j].y = tex.p[3 * (it - 1) + 1]; }
          if (in >= 1 && in <= nrm.n / 3) { t->normals[j].x = nrm.p[3 * (in - 1)]; t->normals[j].y = nrm.p[3 * (in - 1) + 1]; t->normals[j].z = nrm.p[3 * (in - 1) + 2]; }
        }
        tri_mtl[n_tris++] = cur_mtl;
      }
    }
  }
  #undef PUSH3

  /* materials: one per MTL entry plus a trailing default */
  model->n_materials = mtls.n + 1;
  model->materials = calloc((size_t)model->n_materials, sizeof(PBR_Shader_Data));
  model->images = calloc((size_t)(mtls.n * 5 + 1), sizeof(Image));
  model->n_images = 0;
  Tex_Ref *refs = NULL; isize n_refs = 0;
  for (isize i = 0; i < mtls.n; i++) {
    Mtl *m = &mtls.items[i];
    PBR_Shader_Data d;
    memset(&d, 0, sizeof d);
    d.base_color.x = m->kd[0]; d.base_color.y = m->kd[1]; d.base_color.z = m->kd[2];
    d.emission.x = m->ke[0]; d.emission.y = m->ke[1]; d.emission.z = m->ke[2];
    d.roughness = 0.5f;
    d.texture_albedo   = texture_for(model, &refs, &n_refs, path, m->map_kd);
    d.texture_emission = texture_for(model, &refs, &n_refs, path, m->map_ke);
    if (m->is_pbr) {
      d.anisotropic_strength    = m->aniso;
      d.metalness               = m->pm;
      d.roughness               = m->pr;
      d.sheen                   = m->ps;
      d.texture_normal          = texture_for(model, &refs, &n_refs, path, m->map_norm);
      d.texture_metal_roughness = texture_for(model, &refs, &n_refs, path, m->map_pm);
    } else {
      fprintf(stderr, "material %ld is not a pbr material\n", (long)i);
    }
    model->materials[i] = d;
  }
  model->materials[mtls.n] = default_material();
  for (isize i = 0; i < n_tris; i++) {
    isize m = tri_mtl[i] >= 0 ? tri_mtl[i] : mtls.n;
    tris[i].shader.data = &model->materials[m];
    tris[i].shader.proc = proc;
  }
  model->triangles.data = tris;
  model->triangles.len = n_tris;
  free(refs); free(tri_mtl); free(mtls.items); free(pos.p); free(tex.p); free(nrm.p);
  return true;
}

/* ==================================================================== glTF */
typedef struct { f64 m[4][4]; } Mat4d;

static Mat4d mat_identity(void) { Mat4d r; memset(&r, 0, sizeof r); for (int i = 0; i < 4; i++) r.m[i][i] = 1; return r; }
static Mat4d mat_mul(Mat4d a, Mat4d b) {
  Mat4d r;
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) {
    f64 s = 0;
    for (int k = 0; k < 4; k++) s += a.m[i][k] * b.m[k][j];
    r.m[i][j] = s;
  }
  return r;
}

static Mat4d node_local(RT_Json const *node) {
  RT_Json *mj = rt_json_get(node, "matrix");
  Mat4d r = mat_identity();
  if (mj && rt_json_len(mj) == 16) {          /* column-major in the file */
    for (int c = 0; c < 4; c++) for (int row = 0; row < 4; row++) r.m[row][c] = rt_json_num(rt_json_at(mj, c * 4 + row), 0);
    return r;
  }
  f64 t[3] = { 0, 0, 0 }, q[4] = { 0, 0, 0, 1 }, s[3] = { 1, 1, 1 };
  RT_Json *tj = rt_json_get(node, "translation"), *qj = rt_json_get(node, "rotation"), *sj = rt_json_get(node, "scale");
  for (int i = 0; i < 3 && tj; i++) t[i] = rt_json_num(rt_json_at(tj, i), 0);
  for (int i = 0; i < 4 && qj; i++) q[i] = rt_json_num(rt_json_at(qj, i), i == 3);
  for (int i = 0; i < 3 && sj; i++) s[i] = rt_json_num(rt_json_at(sj, i), 1);
  f64 x = q[0], y = q[1], z = q[2], w = q[3];
  f64 rot[3][3] = {
    { 1 - 2 * (y * y + z * z), 2 * (x * y - z * w),     2 * (x * z + y * w) },
    { 2 * (x * y + z * w),     1 - 2 * (x * x + z * z), 2 * (y * z - x * w) },
    { 2 * (x * z - y * w),     2 * (y * z + x * w),     1 - 2 * (x * x + y * y) } };
  for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) r.m[i][j] = rot[i][j] * s[j]; r.m[i][3] = t[i]; }
  return r;
}

typedef struct {
  RT_Json  *root;
  u8      **buffers; size_t *buffer_len; isize n_buffers;
} Gltf;

typedef struct { u8 const *base; isize count, stride, comp_type, n_comp; bool normalized; } Accessor;

static isize comp_size(isize t) { return (t == 5120 || t == 5121) ? 1 : (t == 5122 || t == 5123) ? 2 : 4; }

static bool accessor_get(Gltf const *g, isize index, Accessor *a) {
  RT_Json *acc = rt_json_at(rt_json_get(g->root, "accessors"), index);
  if (!acc) return false;
  RT_Json *bv = rt_json_at(rt_json_get(g->root, "bufferViews"), rt_json_int(rt_json_get(acc, "bufferView"), -1));
  if (!bv) return false;
  isize buf = rt_json_int(rt_json_get(bv, "buffer"), 0);
  if (buf < 0 || buf >= g->n_buffers || !g->buffers[buf]) return false;
  char const *type = rt_json_str(rt_json_get(acc, "type"));
  a->n_comp = !type ? 1 : !strcmp(type, "VEC2") ? 2 : !strcmp(type, "VEC3") ? 3 : !strcmp(type, "VEC4") ? 4 : !strcmp(type, "MAT4") ? 16 : 1;
  a->comp_type = rt_json_int(rt_json_get(acc, "componentType"), 5126);
  a->count = rt_json_int(rt_json_get(acc, "count"), 0);
  a->normalized = rt_json_num(rt_json_get(acc, "normalized"), 0) != 0;
  isize stride = rt_json_int(rt_json_get(bv, "byteStride"), 0);
  isize elem = a->n_comp * comp_size(a->comp_type);
  a->stride = stride ? stride : elem;
  /* every number below comes from the file: reject negatives and check ranges without letting a sum wrap */
  isize off_view = rt_json_int(rt_json_get(bv, "byteOffset"), 0), off_acc = rt_json_int(rt_json_get(acc, "byteOffset"), 0);
  if (a->count < 0 || a->stride < elem || a->stride > (1 << 20) || off_view < 0 || off_acc < 0) return false;
  size_t buflen = g->buffer_len[buf];
  if ((size_t)off_view > buflen || (size_t)off_acc > buflen - (size_t)off_view) return false;
  size_t off = (size_t)off_view + (size_t)off_acc, room = buflen - off;
  if (a->count > 0) {
    if ((size_t)elem > room) return false;
    if ((size_t)(a->count - 1) > (room - (size_t)elem) / (size_t)a->stride) return false;
  }
  a->base = g->buffers[buf] + off;
  return true;
}

static f32 accessor_f32(Accessor const *a, isize i, isize c) {
  u8 const *p = a->base + i * a->stride + c * comp_size(a->comp_type);
  switch (a->comp_type) {
    case 5126: { f32 v; memcpy(&v, p, 4); return v; }
    case 5121: return a->normalized ? p[0] / 255.0f : (f32)p[0];
    case 5123: { unsigned short v; memcpy(&v, p, 2); return a->normalized ? v / 65535.0f : (f32)v; }
    case 5120: { signed char v = (signed char)p[0]; return a->normalized ? fmaxf(v / 127.0f, -1.0f) : (f32)v; }
    case 5122: { short v; memcpy(&v, p, 2); return a->normalized ? fmaxf(v / 32767.0f, -1.0f) : (f32)v; }
    default:   { u32 v; memcpy(&v, p, 4); return (f32)v; }
  }
}

static u32 accessor_u32(Accessor const *a, isize i) {
  u8 const *p = a->base + i * a->stride;
  switch (a->comp_type) {
    case 5121: return p[0];
    case 5123: { unsigned short v; memcpy(&v, p, 2); return v; }
    default:   { u32 v; memcpy(&v, p, 4); return v; }
  }
}

static isize texture_image(Gltf const *g, RT_Json const *tex_info) {
  if (!tex_info) return -1;
  RT_Json *tex = rt_json_at(rt_json_get(g->root, "textures"), rt_json_int(rt_json_get(tex_info, "index"), -1));
  return tex ? rt_json_int(rt_json_get(tex, "source"), -1) : -1;
}

typedef struct { Triangle *t; isize n, cap; } Tri_Vec;

static void emit_mesh(Gltf const *g, RT_Json const *mesh, Mat4d xf, RT_Model *model, Shader_Proc proc, Tri_Vec *out) {
  /* normal matrix = inverse transpose of the linear part = cofactor matrix / det */
  f64 a[3][3], cof[3][3];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) a[i][j] = xf.m[i][j];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
    int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
    cof[i][j] = a[i1][j1] * a[i2][j2] - a[i1][j2] * a[i2][j1];
  }
  RT_Json *prims = rt_json_get(mesh, "primitives");
  for (RT_Json *prim = prims ? prims->first : NULL; prim; prim = prim->next) {
    if (rt_json_int(rt_json_get(prim, "mode"), 4) != 4) continue;
    RT_Json *attrs = rt_json_get(prim, "attributes");
    Accessor P, N, T, I;
    bool has_p = accessor_get(g, rt_json_int(rt_json_get(attrs, "POSITION"), -1), &P);
    bool has_n = accessor_get(g, rt_json_int(rt_json_get(attrs, "NORMAL"), -1), &N);
    bool has_t = accessor_get(g, rt_json_int(rt_json_get(attrs, "TEXCOORD_0"), -1), &T);
    bool has_i = accessor_get(g, rt_json_int(rt_json_get(prim, "indices"), -1), &I);
    if (!has_p || P.n_comp < 3 || P.count < 1) continue;
    if (has_n && N.n_comp < 3) has_n = false;
    if (has_t && T.n_comp < 2) has_t = false;
    isize material = rt_json_int(rt_json_get(prim, "material"), -1);
    if (material < 0 || material >= model->n_materials - 1) material = model->n_materials - 1;
    isize n_idx = has_i ? I.count : P.count;
    for (isize k = 0; k + 2 < n_idx; k += 3) {
      if (out->n == out->cap) {
        isize cap = out->cap ? out->cap * 2 : 4096;
        Triangle *grown = realloc(out->t, sizeof(Triangle) * (size_t)cap);
        if (!grown) { rt_host_set_error("gltf: out of memory"); return; }
        out->t = grown; out->cap = cap;
      }
      Triangle *t = &out->t[out->n++];
      memset(t, 0, sizeof *t);
      for (int j = 0; j < 3; j++) {
        isize v = has_i ? (isize)accessor_u32(&I, k + j) : k + j;
        if (v >= P.count) v = 0;
        f64 p[3] = { accessor_f32(&P, v, 0), accessor_f32(&P, v, 1), accessor_f32(&P, v, 2) };
        for (int r = 0; r < 3; r++) t->positions[j].data[r] = (f32)(xf.m[r][0] * p[0] + xf.m[r][1] * p[1] + xf.m[r][2] * p[2] + xf.m[r][3]);
        if (has_n && v < N.count) {
          f64 n[3] = { accessor_f32(&N, v, 0), accessor_f32(&N, v, 1), accessor_f32(&N, v, 2) };
          f64 w[3], len = 0;
          for (int r = 0; r < 3; r++) { w[r] = cof[r][0] * n[0] + cof[r][1] * n[1] + cof[r][2] * n[2]; len += w[r] * w[r]; }
          len = len > 0 ? 1.0 / sqrt(len) : 0;
          for (int r = 0; r < 3; r++) t->normals[j].data[r] = (f32)(w[r] * len);
        }
        if (has_t && v < T.count) { t->tex_coords[j].x = accessor_f32(&T, v, 0); t->tex_coords[j].y = accessor_f32(&T, v, 1); }
      }
      t->shader.data = &model->materials[material];
      t->shader.proc = proc;
    }
  }
}

static bool load_gltf(char const *path, u8 *data, size_t len, Shader_Proc proc, RT_Model *model, Camera *camera) {
  Gltf g;
  memset(&g, 0, sizeof g);
  char const *json = (char const *)data; size_t json_len = len;
  u8 const *bin = NULL; size_t bin_len = 0;
  if (len >= 20 && !memcmp(data, "glTF", 4)) {
    u32 clen; memcpy(&clen, data + 12, 4);
    if ((size_t)clen > len - 20) { rt_host_set_error("glb: JSON chunk length exceeds the file"); return false; }
    json = (char const *)data + 20; json_len = clen;
    size_t off = 20 + (size_t)clen;
    if (off + 8 <= len) { u32 blen; memcpy(&blen, data + off, 4); bin = data + off + 8; bin_len = blen; if (bin_len > len - off - 8) bin_len = len - off - 8; }
  }
  RT_Json_Doc *doc = NULL;
  g.root = rt_json_parse(json, json_len, &doc);
  if (!g.root) { rt_host_set_error("gltf: JSON parse failed"); return false; }

  /* buffers (driver.c:596 gltf_load_buffers) */
  RT_Json *bufs = rt_json_get(g.root, "buffers");
  g.n_buffers = rt_json_len(bufs);
  g.buffers = calloc((size_t)g.n_buffers + 1, sizeof(u8 *));
  g.buffer_len = calloc((size_t)g.n_buffers + 1, sizeof(size_t));
  bool *owned = calloc((size_t)g.n_buffers + 1, sizeof(bool));
  if (!g.buffers || !g.buffer_len || !owned) { rt_host_set_error("gltf: out of memory"); free(g.buffers); free(g.buffer_len); free(owned); rt_json_free(doc); return false; }
  for (isize i = 0; i < g.n_buffers; i++) {
    char const *uri = rt_json_str(rt_json_get(rt_json_at(bufs, i), "uri"));
    if (!uri) { g.buffers[i] = (u8 *)bin; g.buffer_len[i] = bin_len; continue; }
    if (!strncmp(uri, "data:", 5)) continue;
    char *full = sibling_path(path, uri);
    g.buffers[i] = slurp(full, &g.buffer_len[i]);
    owned[i] = true;
    free(full);
  }

  /* camera: first perspective camera node (driver.c:599-612), global transform */
  RT_Json *nodes = rt_json_get(g.root, "nodes");
  isize n_nodes = rt_json_len(nodes);
  Mat4d *global = malloc(sizeof(Mat4d) * (size_t)(n_nodes + 1));
  isize *parent = malloc(sizeof(isize) * (size_t)(n_nodes + 1));
  if (!global || !parent) {
    rt_host_set_error("gltf: out of memory");
    for (isize i = 0; i < g.n_buffers; i++) if (owned[i]) free(g.buffers[i]);
    free(owned); free(g.buffers); free(g.buffer_len); free(global); free(parent); rt_json_free(doc);
    return false;
  }
  for (isize i = 0; i < n_nodes; i++) parent[i] = -1;
  for (isize i = 0; i < n_nodes; i++) {
    RT_Json *kids = rt_json_get(rt_json_at(nodes, i), "children");
    for (RT_Json *k = kids ? kids->first : NULL; k; k = k->next) { isize c = rt_json_int(k, -1); if (c >= 0 && c < n_nodes) parent[c] = i; }
  }
  for (isize i = 0; i < n_nodes; i++) {
    Mat4d m = node_local(rt_json_at(nodes, i));
    for (isize p = parent[i], guard = 0; p >= 0 && guard < 256; p = parent[p], guard++) m = mat_mul(node_local(rt_json_at(nodes, p)), m);
    global[i] = m;
  }
  for (isize i = 0; i < n_nodes; i++) {
    RT_Json *node = rt_json_at(nodes, i);
    isize cam = rt_json_int(rt_json_get(node, "camera"), -1);
    if (cam < 0) continue;
    RT_Json *cj = rt_json_at(rt_json_get(g.root, "cameras"), cam);
    RT_Json *persp = rt_json_get(cj, "perspective");
    if (!persp) continue;                       /* orthographic: skipped, driver.c:602-604 */
    f32 yfov = (f32)rt_json_num(rt_json_get(persp, "yfov"), 1.0);
    camera->fov = yfov;
    camera->focal_length = 1.0f / tanf(yfov * 0.5f);
    for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) camera->view_matrix.rows[r][c] = (f32)global[i].m[r][c];
    break;
  }

  /* images (driver.c:617-626) */
  RT_Json *images = rt_json_get(g.root, "images");
  model->n_images = rt_json_len(images);
  model->images = calloc((size_t)model->n_images + 1, sizeof(Image));
  if (!model->images) { rt_host_set_error("gltf: out of memory"); model->n_images = 0; }
  /* The reference decodes the images one after the other (stb_image, driver.c:620-626); decoding
   * four 2048x2048 JPEGs is the largest term of time-to-image once the render runs on GPUs, so each
   * image gets its own thread here.  Results do not depend on the schedule. */
  bool ok = true;
  Decode_Job *jobs = calloc((size_t)model->n_images + 1, sizeof(Decode_Job));
  if (!jobs) { rt_host_set_error("gltf: out of memory"); model->n_images = 0; ok = false; }
  for (isize i = 0; i < model->n_images; i++) {
    RT_Json *im = rt_json_at(images, i);
    isize bv_index = rt_json_int(rt_json_get(im, "bufferView"), -1);
    char const *uri = rt_json_str(rt_json_get(im, "uri"));
    jobs[i].out = &model->images[i];
    if (bv_index >= 0) {
      RT_Json *bv = rt_json_at(rt_json_get(g.root, "bufferViews"), bv_index);
      isize buf = rt_json_int(rt_json_get(bv, "buffer"), 0);
      isize off = rt_json_int(rt_json_get(bv, "byteOffset"), 0), n = rt_json_int(rt_json_get(bv, "byteLength"), 0);
      if (bv && buf >= 0 && buf < g.n_buffers && g.buffers[buf] && off >= 0 && n >= 0 && (size_t)off <= g.buffer_len[buf] &&
          (size_t)n <= g.buffer_len[buf] - (size_t)off) { jobs[i].bytes = g.buffers[buf] + off; jobs[i].len = (size_t)n; }
      else snprintf(jobs[i].error, sizeof jobs[i].error, "bufferView %ld is out of range", (long)bv_index);
    } else if (uri && strncmp(uri, "data:", 5)) {
      jobs[i].path = sibling_path(path, uri);
    } else snprintf(jobs[i].error, sizeof jobs[i].error, "image has neither a bufferView nor a file uri");
  }
  for (isize i = 1; i < model->n_images; i++) jobs[i].threaded = pthread_create(&jobs[i].thread, NULL, decode_job_run, &jobs[i]) == 0;
  for (isize i = 0; i < model->n_images; i++) {
    if (jobs[i].threaded) pthread_join(jobs[i].thread, NULL);
    else decode_job_run(&jobs[i]);
    if (!jobs[i].ok) {
      ok = false;
      fprintf(stderr, "Failed to load image %ld of '%s': %s\n", (long)i, path, jobs[i].error);
    }
    free(jobs[i].path);
  }
  free(jobs);

  /* materials (driver.c:628-660); one extra default at the end */
  RT_Json *mats = rt_json_get(g.root, "materials");
  isize n_mats = rt_json_len(mats);
  model->n_materials = n_mats + 1;
  model->materials = calloc((size_t)model->n_materials, sizeof(PBR_Shader_Data));
  for (isize i = 0; i < n_mats && ok; i++) {
    RT_Json *m = rt_json_at(mats, i), *pbr = rt_json_get(m, "pbrMetallicRoughness");
    PBR_Shader_Data d;
    memset(&d, 0, sizeof d);
    RT_Json *bc = rt_json_get(pbr, "baseColorFactor"), *em = rt_json_get(m, "emissiveFactor");
    for (int c = 0; c < 3; c++) {
      d.base_color.data[c] = (f32)rt_json_num(rt_json_at(bc, c), 1.0);
      d.emission.data[c]   = (f32)rt_json_num(rt_json_at(em, c), 0.0);
    }
    d.roughness = (f32)rt_json_num(rt_json_get(pbr, "roughnessFactor"), 1.0);
    d.metalness = (f32)rt_json_num(rt_json_get(pbr, "metallicFactor"), 1.0);
    RT_Json *sheen = rt_json_get(rt_json_get(m, "extensions"), "KHR_materials_sheen");
    if (sheen) {
      RT_Json *sc = rt_json_get(sheen, "sheenColorFactor");
      f32 c3[3] = { (f32)rt_json_num(rt_json_at(sc, 0), 0), (f32)rt_json_num(rt_json_at(sc, 1), 0), (f32)rt_json_num(rt_json_at(sc, 2), 0) };
      d.sheen = c3[0] * 0.2126f + c3[1] * 0.7152f + c3[2] * 0.0722f;     /* driver.c:637 */
    }
    isize img;
    RT_Json *nt = rt_json_get(m, "normalTexture");
    if ((img = texture_image(&g, nt)) >= 0 && img < model->n_images) {
      d.texture_normal = &model->images[img];
      d.normal_map_strength = (f32)rt_json_num(rt_json_get(nt, "scale"), 1.0);
    }
    if ((img = texture_image(&g, rt_json_get(m, "emissiveTexture"))) >= 0 && img < model->n_images) d.texture_emission = &model->images[img];
    if ((img = texture_image(&g, rt_json_get(pbr, "baseColorTexture"))) >= 0 && img < model->n_images) d.texture_albedo = &model->images[img];
    if ((img = texture_image(&g, rt_json_get(pbr, "metallicRoughnessTexture"))) >= 0 && img < model->n_images) d.texture_metal_roughness = &model->images[img];
    model->materials[i] = d;
  }
  model->materials[n_mats] = default_material();

  /* triangles (driver.c:662-681) */
  Tri_Vec tv = {0};
  if (ok) {
    RT_Json *meshes = rt_json_get(g.root, "meshes");
    for (isize i = 0; i < n_nodes; i++) {
      isize mesh = rt_json_int(rt_json_get(rt_json_at(nodes, i), "mesh"), -1);
      if (mesh >= 0) emit_mesh(&g, rt_json_at(meshes, mesh), global[i], model, proc, &tv);
    }
  }
  model->triangles.data = tv.t;
  model->triangles.len = tv.n;

  for (isize i = 0; i < g.n_buffers; i++) if (owned[i]) free(g.buffers[i]);
  free(owned); free(g.buffers); free(g.buffer_len); free(global); free(parent);
  rt_json_free(doc);
  return ok;
}

/* driver.c:685-728 */
bool rt_load_model_file(char const *path, Shader_Proc proc, RT_Model *model, Camera *camera) {
  memset(model, 0, sizeof *model);
  char const *ext = strrchr(path, '.');
  if (!ext || (strcmp(ext, ".obj") && strcmp(ext, ".glb") && strcmp(ext, ".gltf"))) {
    fprintf(stderr, "Unrecognized file type: '%s'\n", path);
    rt_host_set_error("model: unrecognized file type");
    return false;
  }
  size_t len;
  u8 *data = slurp(path, &len);
  if (!data) { fprintf(stderr, "Failed to read model file\n"); rt_host_set_error("model: cannot read file"); return false; }
  bool ok = !strcmp(ext, ".obj") ? load_obj(path, data, proc, model) : load_gltf(path, data, len, proc, model, camera);
  free(data);
  if (!ok) rt_model_free(model);
  return ok;
}

void rt_model_free(RT_Model *model) {
  free(model->triangles.data);
  for (isize i = 0; i < model->n_images; i++) rt_host_buffer_free(model->images[i].pixels.data);
  free(model->images);
  free(model->materials);
  memset(model, 0, sizeof *model);
}
