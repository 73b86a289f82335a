#include "json_min.h"

#include <ctype.h>
#include <errno.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define JSON_MAX_DEPTH 64

struct parser { const char *p; char *err; size_t errlen; int failed; int depth; };

static void fail(struct parser *ps, const char *what)
{
    if (!ps->failed && ps->err) snprintf(ps->err, ps->errlen, "JSON: %s near '%.20s'", what, ps->p);
    ps->failed = 1;
}

static void skip_ws(struct parser *ps) { while (*ps->p && isspace((unsigned char)*ps->p)) ps->p++; }

static jnode *node_new(enum jtype t)
{
    jnode *n = calloc(1, sizeof(*n));
    if (n) n->type = t;
    return n;
}

void json_free(jnode *n)
{
    if (!n) return;
    for (size_t i = 0; i < n->len; i++) {
        if (n->items) json_free(n->items[i]);
        if (n->keys) free(n->keys[i]);
    }
    free(n->items); free(n->keys); free(n->str); free(n);
}

static jnode *parse_value(struct parser *ps);

static char *parse_string_raw(struct parser *ps)
{
    if (*ps->p != '"') { fail(ps, "expected string"); return NULL; }
    ps->p++;
    size_t cap = 32, len = 0;
    char *out = malloc(cap);
    if (!out) { fail(ps, "out of memory"); return NULL; }
    while (*ps->p && *ps->p != '"') {
        char ch = *ps->p++;
        if (ch == '\\') {
            char e = *ps->p;
            if (e == 0) { free(out); fail(ps, "unterminated string"); return NULL; }     /* text ends in a lone backslash */
            ps->p++;
            switch (e) {
            case 'n': ch = '\n'; break; case 't': ch = '\t'; break; case 'r': ch = '\r'; break;
            case 'b': ch = '\b'; break; case 'f': ch = '\f'; break;
            case 'u': {
                unsigned v = 0;
                for (int i = 0; i < 4 && isxdigit((unsigned char)*ps->p); i++) {
                    char h = *ps->p++;
                    v = v * 16 + (unsigned)(isdigit((unsigned char)h) ? h - '0' : (tolower(h) - 'a' + 10));
                }
                ch = (char)(v < 128 ? v : '?');
                break;
            }
            default: ch = e; break;
            }
        }
        if (len + 2 > cap) {
            char *grown = realloc(out, cap * 2);
            if (!grown) { free(out); fail(ps, "out of memory"); return NULL; }
            out = grown; cap *= 2;
        }
        out[len++] = ch;
    }
    if (*ps->p != '"') { free(out); fail(ps, "unterminated string"); return NULL; }
    ps->p++;
    out[len] = 0;
    return out;
}

/* 0 on success; on failure nothing was attached (the caller still owns key and v) */
static int push(jnode *n, char *key, jnode *v)
{
    jnode **items = realloc(n->items, (n->len + 1) * sizeof(*n->items));
    if (!items) return -1;
    n->items = items;
    if (n->type == J_OBJ) {
        char **keys = realloc(n->keys, (n->len + 1) * sizeof(*n->keys));
        if (!keys) return -1;
        n->keys = keys;
        n->keys[n->len] = key;
    }
    n->items[n->len++] = v;
    return 0;
}

static jnode *parse_value_inner(struct parser *ps);

static jnode *parse_value(struct parser *ps)
{
    if (ps->depth >= JSON_MAX_DEPTH) { fail(ps, "nesting too deep"); return NULL; }
    ps->depth++;
    jnode *n = parse_value_inner(ps);
    ps->depth--;
    if (!n && !ps->failed) fail(ps, "out of memory");
    return n;
}

static jnode *parse_value_inner(struct parser *ps)
{
    skip_ws(ps);
    const char c = *ps->p;
    if (c == '{') {
        jnode *n = node_new(J_OBJ);
        if (!n) return NULL;
        ps->p++; skip_ws(ps);
        if (*ps->p == '}') { ps->p++; return n; }
        for (;;) {
            skip_ws(ps);
            char *k = parse_string_raw(ps);
            if (!k) { json_free(n); return NULL; }
            skip_ws(ps);
            if (*ps->p != ':') { free(k); json_free(n); fail(ps, "expected ':'"); return NULL; }
            ps->p++;
            jnode *v = parse_value(ps);
            if (!v) { free(k); json_free(n); return NULL; }
            if (push(n, k, v)) { free(k); json_free(v); json_free(n); fail(ps, "out of memory"); return NULL; }
            skip_ws(ps);
            if (*ps->p == ',') { ps->p++; continue; }
            if (*ps->p == '}') { ps->p++; return n; }
            json_free(n); fail(ps, "expected ',' or '}'"); return NULL;
        }
    }
    if (c == '[') {
        jnode *n = node_new(J_ARR);
        if (!n) return NULL;
        ps->p++; skip_ws(ps);
        if (*ps->p == ']') { ps->p++; return n; }
        for (;;) {
            jnode *v = parse_value(ps);
            if (!v) { json_free(n); return NULL; }
            if (push(n, NULL, v)) { json_free(v); json_free(n); fail(ps, "out of memory"); return NULL; }
            skip_ws(ps);
            if (*ps->p == ',') { ps->p++; continue; }
            if (*ps->p == ']') { ps->p++; return n; }
            json_free(n); fail(ps, "expected ',' or ']'"); return NULL;
        }
    }
    if (c == '"') {
        char *s = parse_string_raw(ps);
        if (!s) return NULL;
        jnode *n = node_new(J_STR);
        if (!n) { free(s); return NULL; }
        n->str = s;
        return n;
    }
    if (!strncmp(ps->p, "true", 4))  { ps->p += 4; jnode *n = node_new(J_BOOL); if (n) n->num = 1; return n; }
    if (!strncmp(ps->p, "false", 5)) { ps->p += 5; jnode *n = node_new(J_BOOL); if (n) n->num = 0; return n; }
    if (!strncmp(ps->p, "null", 4))  { ps->p += 4; return node_new(J_NULL); }
    if (c == '-' || c == '+' || isdigit((unsigned char)c) || c == '.') {
        char *end = NULL;
        errno = 0;
        double v = strtod(ps->p, &end);
        if (end == ps->p) { fail(ps, "bad number"); return NULL; }
        jnode *n = node_new(J_NUM);
        if (!n) return NULL;
        n->num = v;
        n->is_int = 1;
        for (const char *q = ps->p; q < end; q++) if (*q == '.' || *q == 'e' || *q == 'E') n->is_int = 0;
        if (n->is_int) {            /* integers are kept exactly as 64-bit values (jansson's json_int_t) */
            errno = 0;
            const long long iv = strtoll(ps->p, NULL, 10);
            if (errno == ERANGE) n->is_int = 0; else n->inum = iv;
        }
        ps->p = end;
        return n;
    }
    fail(ps, "unexpected character");
    return NULL;
}

jnode *json_parse_text(const char *text, char *err, size_t errlen)
{
    struct parser ps = { text, err, errlen, 0, 0 };
    jnode *n = parse_value(&ps);
    if (n) {
        skip_ws(&ps);
        if (*ps.p) { fail(&ps, "trailing data"); json_free(n); return NULL; }
    }
    return n;
}

jnode *json_parse_file(const char *path, char *err, size_t errlen)
{
    FILE *f = fopen(path, "rb");
    if (!f) { if (err) snprintf(err, errlen, "cannot open %s: %s", path, strerror(errno)); return NULL; }
    long sz = -1;
    if (0 == fseek(f, 0, SEEK_END)) sz = ftell(f);
    if (sz < 0 || fseek(f, 0, SEEK_SET)) {
        if (err) snprintf(err, errlen, "cannot size %s: %s", path, strerror(errno));
        fclose(f);
        return NULL;
    }
    char *buf = malloc((size_t)sz + 1);
    if (!buf) { if (err) snprintf(err, errlen, "out of memory reading %s", path); fclose(f); return NULL; }
    size_t rd = fread(buf, 1, (size_t)sz, f);
    fclose(f);
    buf[rd] = 0;
    jnode *n = json_parse_text(buf, err, errlen);
    free(buf);
    return n;
}

int json_merge(jnode *dst, jnode *src)
{
    if (!dst || !src || dst->type != J_OBJ || src->type != J_OBJ) return -1;
    for (size_t i = 0; i < src->len; i++) {
        size_t j;
        for (j = 0; j < dst->len; j++) if (!strcmp(dst->keys[j], src->keys[i])) break;
        if (j < dst->len) { json_free(dst->items[j]); dst->items[j] = src->items[i]; free(src->keys[i]); }
        else if (push(dst, src->keys[i], src->items[i])) return -1;
        src->items[i] = NULL; src->keys[i] = NULL;
    }
    src->len = 0;
    json_free(src);
    return 0;
}

const jnode *json_get(const jnode *obj, const char *key)
{
    if (!obj || obj->type != J_OBJ) return NULL;
    for (size_t i = 0; i < obj->len; i++) if (!strcmp(obj->keys[i], key)) return obj->items[i];
    return NULL;
}

/* The reference reads integers through the TSL config engine on top of jansson: a 64-bit json_int_t stored into an
 * `int` (multifm/receiver.c:139-160, 204), i.e. truncated modulo 2^32.  A 2.4 GHz centerFreqHz therefore wraps to a
 * negative int and `(int32_t)nb_center_freq - center_freq` (receiver.c:229) still yields the right offset.  Same
 * here: integers only (a number with a fraction or an exponent is a type error), wrapped like the C conversion. */
int json_get_int(const jnode *obj, const char *key, int *out)
{
    const jnode *n = json_get(obj, key);
    if (!n || n->type != J_NUM || !n->is_int) return -1;
    *out = (int)(int32_t)(uint32_t)(uint64_t)n->inum;
    return 0;
}

int json_get_int64(const jnode *obj, const char *key, long long *out)
{
    const jnode *n = json_get(obj, key);
    if (!n || n->type != J_NUM || !n->is_int) return -1;
    *out = n->inum;
    return 0;
}

int json_get_double(const jnode *obj, const char *key, double *out)
{
    const jnode *n = json_get(obj, key);
    if (!n || n->type != J_NUM) return -1;
    *out = n->num;
    return 0;
}

int json_get_string(const jnode *obj, const char *key, const char **out)
{
    const jnode *n = json_get(obj, key);
    if (!n || n->type != J_STR) return -1;
    *out = n->str;
    return 0;
}
