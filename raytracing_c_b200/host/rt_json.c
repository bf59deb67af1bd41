/* rt_json.c — small JSON DOM for the glTF loader (stands in for Codin's
 * gltf_parse, reference driver.c:594).  One arena per document. */
#include "rt_json.h"

#include <stdlib.h>
#include <string.h>

typedef struct Block { struct Block *next; size_t used, cap; } Block;

struct RT_Json_Doc { Block *blocks; char const *p, *end; bool failed; };

static void *arena_alloc(RT_Json_Doc *d, size_t n) {
  n = (n + 15) & ~(size_t)15;
  if (!d->blocks || d->blocks->used + n > d->blocks->cap) {
    size_t cap = n > (1u << 16) ? n : (1u << 16);
    Block *b = malloc(sizeof(Block) + cap);
    b->next = d->blocks; b->used = 0; b->cap = cap;
    d->blocks = b;
  }
  void *out = (char *)(d->blocks + 1) + d->blocks->used;
  d->blocks->used += n;
  memset(out, 0, n);
  return out;
}

static void skip_ws(RT_Json_Doc *d) {
  while (d->p < d->end && (*d->p == ' ' || *d->p == '\t' || *d->p == '\n' || *d->p == '\r')) d->p++;
}

static RT_Json *parse_value(RT_Json_Doc *d, int depth);

static char *parse_string_raw(RT_Json_Doc *d) {
  if (d->p >= d->end || *d->p != '"') { d->failed = true; return NULL; }
  d->p++;
  char const *start = d->p;
  size_t n = 0;
  while (d->p < d->end && *d->p != '"') { if (*d->p == '\\') d->p++; d->p++; n++; }
  if (d->p >= d->end) { d->failed = true; return NULL; }
  char *out = arena_alloc(d, n + 1), *w = out;
  for (char const *r = start; r < d->p; r++) {
    if (*r != '\\') { *w++ = *r; continue; }
    r++;
    switch (*r) {
      case 'n': *w++ = '\n'; break;
      case 't': *w++ = '\t'; break;
      case 'r': *w++ = '\r'; break;
      case 'b': *w++ = '\b'; break;
      case 'f': *w++ = '\f'; break;
      case 'u': *w++ = '?'; r += 4; break;   /* names only; not needed for paths here */
      default:  *w++ = *r; break;
    }
  }
  *w = 0;
  d->p++;
  return out;
}

static RT_Json *parse_value(RT_Json_Doc *d, int depth) {
  if (depth > 64) { d->failed = true; return NULL; }
  skip_ws(d);
  if (d->p >= d->end) { d->failed = true; return NULL; }
  RT_Json *v = arena_alloc(d, sizeof *v);
  char c = *d->p;
  if (c == '{' || c == '[') {
    bool is_obj = c == '{';
    v->type = is_obj ? RT_JSON_OBJECT : RT_JSON_ARRAY;
    d->p++;
    RT_Json **tail = &v->first;
    skip_ws(d);
    if (d->p < d->end && *d->p == (is_obj ? '}' : ']')) { d->p++; return v; }
    for (;;) {
      char *key = NULL;
      if (is_obj) {
        skip_ws(d);
        key = parse_string_raw(d);
        skip_ws(d);
        if (d->failed || d->p >= d->end || *d->p != ':') { d->failed = true; return NULL; }
        d->p++;
      }
      RT_Json *child = parse_value(d, depth + 1);
      if (d->failed || !child) { d->failed = true; return NULL; }
      child->key = key;
      *tail = child;
      tail = &child->next;
      v->count++;
      skip_ws(d);
      if (d->p >= d->end) { d->failed = true; return NULL; }
      if (*d->p == ',') { d->p++; continue; }
      if (*d->p == (is_obj ? '}' : ']')) { d->p++; return v; }
      d->failed = true;
      return NULL;
    }
  }
  if (c == '"') { v->type = RT_JSON_STRING; v->string = parse_string_raw(d); return v; }
  if (c == 't' && d->end - d->p >= 4 && !memcmp(d->p, "true", 4))  { v->type = RT_JSON_BOOL; v->number = 1; d->p += 4; return v; }
  if (c == 'f' && d->end - d->p >= 5 && !memcmp(d->p, "false", 5)) { v->type = RT_JSON_BOOL; v->number = 0; d->p += 5; return v; }
  if (c == 'n' && d->end - d->p >= 4 && !memcmp(d->p, "null", 4))  { v->type = RT_JSON_NULL; d->p += 4; return v; }
  {
    char buf[64];
    size_t n = 0;
    while (d->p + n < d->end && n < sizeof buf - 1 && strchr("+-0123456789.eE", d->p[n])) n++;
    if (!n) { d->failed = true; return NULL; }
    memcpy(buf, d->p, n);
    buf[n] = 0;
    v->type = RT_JSON_NUMBER;
    v->number = strtod(buf, NULL);
    d->p += n;
    return v;
  }
}

RT_Json *rt_json_parse(char const *text, size_t len, RT_Json_Doc **doc_out) {
  RT_Json_Doc *d = calloc(1, sizeof *d);
  d->p = text;
  d->end = text + len;
  RT_Json *root = parse_value(d, 0);
  if (d->failed || !root) { rt_json_free(d); return NULL; }
  *doc_out = d;
  return root;
}

void rt_json_free(RT_Json_Doc *d) {
  if (!d) return;
  for (Block *b = d->blocks; b;) { Block *n = b->next; free(b); b = n; }
  free(d);
}

RT_Json *rt_json_get(RT_Json const *obj, char const *key) {
  if (!obj || obj->type != RT_JSON_OBJECT) return NULL;
  for (RT_Json *c = obj->first; c; c = c->next) if (c->key && !strcmp(c->key, key)) return c;
  return NULL;
}

RT_Json *rt_json_at(RT_Json const *arr, isize index) {
  if (!arr || arr->type != RT_JSON_ARRAY) return NULL;
  RT_Json *c = arr->first;
  while (c && index-- > 0) c = c->next;
  return c;
}

isize rt_json_len(RT_Json const *v) { return v ? v->count : 0; }

double rt_json_num(RT_Json const *v, double fallback) {
  return (v && (v->type == RT_JSON_NUMBER || v->type == RT_JSON_BOOL)) ? v->number : fallback;
}

isize rt_json_int(RT_Json const *v, isize fallback) {
  return (v && v->type == RT_JSON_NUMBER) ? (isize)v->number : fallback;
}

char const *rt_json_str(RT_Json const *v) { return (v && v->type == RT_JSON_STRING) ? v->string : NULL; }
