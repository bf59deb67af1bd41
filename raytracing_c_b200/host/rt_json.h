/* rt_json.h — minimal JSON DOM used by the glTF loader. */
#ifndef RT_JSON_H
#define RT_JSON_H

#include "rt_base.h"

typedef enum { RT_JSON_NULL, RT_JSON_BOOL, RT_JSON_NUMBER, RT_JSON_STRING, RT_JSON_ARRAY, RT_JSON_OBJECT } RT_Json_Type;

typedef struct RT_Json {
  RT_Json_Type    type;
  char           *key;      /* set when the value is an object member */
  char           *string;
  double          number;
  struct RT_Json *first, *next;
  isize           count;
} RT_Json;

typedef struct RT_Json_Doc RT_Json_Doc;

RT_Json    *rt_json_parse(char const *text, size_t len, RT_Json_Doc **doc_out);
void        rt_json_free(RT_Json_Doc *doc);
RT_Json    *rt_json_get(RT_Json const *obj, char const *key);
RT_Json    *rt_json_at(RT_Json const *arr, isize index);
isize       rt_json_len(RT_Json const *v);
double      rt_json_num(RT_Json const *v, double fallback);
isize       rt_json_int(RT_Json const *v, isize fallback);
char const *rt_json_str(RT_Json const *v);

#endif
