/* The reference call sequences against the headers under include/compat (test program, compiled by tests/test_compat_*.py):
 *   POCSAG / FLEX : decoder/decoder.c:685-697 (pager_*_new with the three/two callbacks) and :635-651 (one
 *                   pager_*_on_pcm call per block of at most 1024 resampled samples), callbacks printing one line each;
 *   FM            : multifm/demod.c:89 (multifm_fm_demod_process on at most 1024 filtered IQ samples per call).
 * usage: compat_decoder POCSAG|FLEX|FM <input int16 file> <output file> [block samples] */
#include <demod_base.h>
#include <fm_demod.h>
#include <pager_flex.h>
#include <pager_pocsag.h>

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define NR_SAMPLES 1024     /* decoder/decoder.c:49 */

static FILE *out;

static void put_hex(const char *data, size_t len)
{
    for (size_t i = 0; i < len; i++) fprintf(out, "%02x", (unsigned char)data[i]);
    fputc('\n', out);
}

static aresult_t on_num(struct pager_pocsag *p, uint16_t baud, uint32_t cap, const char *data, size_t len, uint8_t fn)
{
    (void)p;
    fprintf(out, "POCSAG NUM %u %u %u %zu ", baud, cap, fn, len);
    put_hex(data, len);
    return A_OK;
}

static aresult_t on_alpha(struct pager_pocsag *p, uint16_t baud, uint32_t cap, const char *data, size_t len, uint8_t fn)
{
    (void)p;
    fprintf(out, "POCSAG ALN %u %u %u %zu ", baud, cap, fn, len);
    put_hex(data, len);
    return A_OK;
}

static aresult_t fx_aln(struct pager_flex *f, uint16_t baud, uint8_t phase, uint8_t cycle, uint8_t frame, uint64_t cap, bool frag,
                        bool maildrop, uint8_t seq, const char *msg, size_t len)
{
    (void)f;
    fprintf(out, "FLEX ALN %u %u %u %u %llu %d %d %u %zu ", baud, phase, cycle, frame, (unsigned long long)cap, frag, maildrop, seq, len);
    put_hex(msg, len);
    return A_OK;
}

static aresult_t fx_num(struct pager_flex *f, uint16_t baud, uint8_t phase, uint8_t cycle, uint8_t frame, uint64_t cap,
                        const char *msg, size_t len)
{
    (void)f;
    fprintf(out, "FLEX NUM %u %u %u %u %llu %zu ", baud, phase, cycle, frame, (unsigned long long)cap, len);
    put_hex(msg, len);
    return A_OK;
}

static aresult_t fx_siv(struct pager_flex *f, uint16_t baud, uint8_t phase, uint8_t cycle, uint8_t frame, uint64_t cap, uint8_t type,
                        uint32_t data)
{
    (void)f;
    fprintf(out, "FLEX SIV %u %u %u %u %llu %u %u\n", baud, phase, cycle, frame, (unsigned long long)cap, type, data);
    return A_OK;
}

int main(int argc, char **argv)
{
    if (argc < 4) { fprintf(stderr, "usage: %s POCSAG|FLEX|FM in out [block]\n", argv[0]); return 2; }
    FILE *in = fopen(argv[2], "rb");
    out = fopen(argv[3], "wb");
    if (!in || !out) { perror("open"); return 2; }
    size_t block = argc > 4 ? (size_t)atol(argv[4]) : NR_SAMPLES;
    if (block == 0 || block > (1u << 20)) block = NR_SAMPLES;
    int16_t *buf = malloc(block * 2 * sizeof(int16_t)), *pcm = malloc(block * sizeof(int16_t));
    aresult_t ret = A_OK;

    if (!strcmp(argv[1], "FM")) {
        struct demod_base *demod = NULL;
        if (FAILED(ret = multifm_fm_demod_init(&demod))) { fprintf(stderr, "multifm_fm_demod_init failed: %d\n", ret); return 1; }
        size_t n;
        while ((n = fread(buf, 2 * sizeof(int16_t), block, in)) > 0) {
            size_t nr_out = 0, nr_bytes = 0;
            if (FAILED(ret = multifm_fm_demod_process(demod, buf, n, pcm, &nr_out, &nr_bytes))) return 1;
            if (nr_out != n || nr_bytes != n * sizeof(int16_t)) return 3;
            fwrite(pcm, 1, nr_bytes, out);
        }
        if (FAILED(multifm_fm_demod_cleanup(&demod)) || demod != NULL) return 3;
    } else if (!strcmp(argv[1], "POCSAG")) {
        struct pager_pocsag *pocsag = NULL;
        if (FAILED(ret = pager_pocsag_new(&pocsag, 152000000u, on_num, on_alpha, false))) { fprintf(stderr, "pager_pocsag_new failed: %d\n", ret); return 1; }
        size_t n;
        while ((n = fread(buf, sizeof(int16_t), block, in)) > 0)
            if (FAILED(ret = pager_pocsag_on_pcm(pocsag, buf, n))) return 1;
        if (A_E_BADARGS != pager_pocsag_on_pcm(pocsag, buf, 0)) return 3;      /* TSL_ASSERT_ARG(0 != nr_samples) */
        if (FAILED(pager_pocsag_delete(&pocsag)) || pocsag != NULL) return 3;
    } else {
        struct pager_flex *flex = NULL;
        if (A_E_BADARGS != pager_flex_new(&flex, 929000000u, fx_aln, NULL, fx_siv)) return 3;   /* on_num_msg is required */
        if (FAILED(ret = pager_flex_new(&flex, 929000000u, fx_aln, fx_num, fx_siv))) { fprintf(stderr, "pager_flex_new failed: %d\n", ret); return 1; }
        size_t n;
        while ((n = fread(buf, sizeof(int16_t), block, in)) > 0)
            if (FAILED(ret = pager_flex_on_pcm(flex, buf, n))) return 1;
        if (FAILED(pager_flex_delete(&flex)) || flex != NULL) return 3;
    }
    fclose(out);
    fclose(in);
    free(buf); free(pcm);
    return 0;
}
