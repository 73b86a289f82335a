/* Tiny JSON reader for the multifm configuration schema.  The reference uses the
 * TSL config engine on top of jansson (multifm/multifm.c:103-111: several files are MERGED); neither is
 * available here, so the host carries its own reader. */
#ifndef B200_JSON_MIN_H
#define B200_JSON_MIN_H
#include <stddef.h>

enum jtype { J_NULL, J_BOOL, J_NUM, J_STR, J_ARR, J_OBJ };

typedef struct jnode {
    enum jtype type;
    double num;             /* J_NUM, J_BOOL */
    int is_int;             /* number had no fraction/exponent and fits 64 bits */
    long long inum;         /* its exact value (is_int) */
    char *str;              /* J_STR */
    struct jnode **items;   /* J_ARR / J_OBJ values */
    char **keys;            /* J_OBJ keys */
    size_t len;
} jnode;

jnode *json_parse_text(const char *text, char *err, size_t errlen);
jnode *json_parse_file(const char *path, char *err, size_t errlen);
void json_free(jnode *n);
/* config_add semantics: top-level keys of src are added to dst, replacing existing keys */
int json_merge(jnode *dst_obj, jnode *src_obj);

const jnode *json_get(const jnode *obj, const char *key);
int json_get_int(const jnode *obj, const char *key, int *out);          /* 0 on success; integers only, wraps modulo 2^32 */
int json_get_int64(const jnode *obj, const char *key, long long *out);
int json_get_double(const jnode *obj, const char *key, double *out);
int json_get_string(const jnode *obj, const char *key, const char **out);
#endif
