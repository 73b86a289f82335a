/* compat.c -- the reference per-object interfaces (the headers under include/compat) over one-channel B200 banks, so that code
 * written against multifm/fm_demod.h, pager/pager_pocsag.h and pager/pager_flex.h (decoder/decoder.c:635-651,685-697;
 * multifm/demod.c:89) links against this library unchanged.  Built as libtslb200_compat.so. */
#include "../../include/compat/demod_base.h"
#include "../../include/compat/fm_demod.h"
#include "../../include/compat/pager_flex.h"
#include "../../include/compat/pager_pocsag.h"
#include "../../include/tslb200_gpuchan.h"
#include "../../include/tslb200_gpupager.h"

#include <stdio.h>
#include <stdlib.h>

#define COMPAT_MAX_FEED (1u << 16)      /* samples per launch; longer blocks are cut up (any chunking decodes the same) */

static aresult_t map_rc(int rc)
{
    switch (rc) {
    case 0: return A_OK;
    case GPUCHAN_E_NOMEM: return A_E_NOMEM;
    case GPUCHAN_E_BADARGS: return A_E_BADARGS;
    case GPUCHAN_E_BUSY: return A_E_BUSY;
    default: return A_E_INVAL;          /* incl. GPUCHAN_E_NODEVICE / GPUCHAN_E_CUDA: there is no CPU path to fall back to */
    }
}

/* ---- multifm/fm_demod.h ---- */
struct multifm_fm_demod {
    struct demod_base demod;            /* first member, like multifm/fm_demod.c:14-18 */
    gpufm_t *fm;
};

aresult_t multifm_fm_demod_init(struct demod_base **pdemod)
{
    if (NULL == pdemod) return A_E_BADARGS;
    *pdemod = NULL;
    struct multifm_fm_demod *d = calloc(1, sizeof(*d));
    if (NULL == d) return A_E_NOMEM;
    const int rc = gpufm_create(&d->fm, 0, 1u << 20, GPUCHAN_F_DEFAULT);
    if (rc) {
        fprintf(stderr, "multifm_fm_demod_init: %s\n", gpuchan_last_error());
        free(d);
        return map_rc(rc);
    }
    *pdemod = (struct demod_base *)d;
    return A_OK;
}

aresult_t multifm_fm_demod_process(struct demod_base *demod, int16_t *in_samples, size_t nr_in_samples,
        int16_t *out_samples, size_t *pnr_out_samples, size_t *pnr_out_bytes)
{
    /* argument checks of multifm/fm_demod.c:47-51 */
    if (NULL == demod || NULL == in_samples || 0 == nr_in_samples || NULL == out_samples || NULL == pnr_out_samples)
        return A_E_BADARGS;
    struct multifm_fm_demod *d = (struct multifm_fm_demod *)demod;
    const int rc = gpufm_process(d->fm, in_samples, nr_in_samples, out_samples);
    if (rc) return map_rc(rc);
    *pnr_out_samples = nr_in_samples;
    if (pnr_out_bytes) *pnr_out_bytes = nr_in_samples * sizeof(int16_t);
    return A_OK;
}

aresult_t multifm_fm_demod_cleanup(struct demod_base **pdemod)
{
    if (NULL == pdemod || NULL == *pdemod) return A_E_BADARGS;
    struct multifm_fm_demod *d = (struct multifm_fm_demod *)*pdemod;
    gpufm_destroy(&d->fm);
    free(d);
    *pdemod = NULL;
    return A_OK;
}

/* ---- shared: a one-channel pager bank fed from host memory, no resampler in front ---- */
static int one_channel_bank(gpupager_t **pbank, uint32_t decoder)
{
    gpupager_cfg pc = { 0 };
    pc.struct_size = sizeof(pc);
    pc.nr_channels = 1;
    pc.device = 0;
    pc.max_feed_samples = COMPAT_MAX_FEED;
    pc.flags = GPUPAGER_F_NO_RESAMPLE;
    pc.decoder = decoder;
    const int rc = gpupager_create(pbank, &pc);
    if (rc) fprintf(stderr, "pager bank: %s\n", gpupager_last_error());
    return rc;
}

/* ---- pager/pager_pocsag.h ---- */
struct pager_pocsag {
    gpupager_t *bank;
    uint32_t freq_hz;
    pager_pocsag_on_numeric_msg_func_t on_numeric;
    pager_pocsag_on_alpha_msg_func_t on_alpha;
    bool skip_bch;
};

static int pocsag_numeric(void *user, uint32_t channel, uint16_t baud, uint32_t capcode, const char *data, size_t len, uint8_t function)
{
    struct pager_pocsag *p = user;
    (void)channel;
    return p->on_numeric ? p->on_numeric(p, baud, capcode, data, len, function) : 0;
}

static int pocsag_alpha(void *user, uint32_t channel, uint16_t baud, uint32_t capcode, const char *data, size_t len, uint8_t function)
{
    struct pager_pocsag *p = user;
    (void)channel;
    return p->on_alpha ? p->on_alpha(p, baud, capcode, data, len, function) : 0;
}

aresult_t pager_pocsag_new(struct pager_pocsag **ppocsag, uint32_t freq_hz, pager_pocsag_on_numeric_msg_func_t on_numeric,
        pager_pocsag_on_alpha_msg_func_t on_alpha, bool skip_bch_decode)
{
    /* pager/pager_pocsag.c:153: only ppocsag is checked (a NULL callback would crash the reference at delivery time;
     * here such messages are dropped) */
    if (NULL == ppocsag) return A_E_BADARGS;
    *ppocsag = NULL;
    struct pager_pocsag *p = calloc(1, sizeof(*p));
    if (NULL == p) return A_E_NOMEM;
    p->freq_hz = freq_hz; p->on_numeric = on_numeric; p->on_alpha = on_alpha; p->skip_bch = skip_bch_decode;
    const int rc = one_channel_bank(&p->bank, GPUPAGER_DECODER_POCSAG);
    if (rc) { free(p); return map_rc(rc); }
    *ppocsag = p;
    return A_OK;
}

aresult_t pager_pocsag_delete(struct pager_pocsag **ppocsag)
{
    if (NULL == ppocsag || NULL == *ppocsag) return A_E_BADARGS;
    gpupager_destroy(&(*ppocsag)->bank);
    free(*ppocsag);
    *ppocsag = NULL;
    return A_OK;
}

aresult_t pager_pocsag_on_pcm(struct pager_pocsag *pocsag, const int16_t *pcm_samples, size_t nr_samples)
{
    if (NULL == pocsag || NULL == pcm_samples || 0 == nr_samples) return A_E_BADARGS;   /* pager_pocsag.c:442-444 */
    for (size_t done = 0; done < nr_samples; ) {
        const size_t n = nr_samples - done < COMPAT_MAX_FEED ? nr_samples - done : COMPAT_MAX_FEED;
        int rc = gpupager_feed(pocsag->bank, pcm_samples + done, n, n);
        if (!rc) rc = gpupager_dispatch(pocsag->bank, pocsag_numeric, pocsag_alpha, pocsag, NULL);
        if (rc) return map_rc(rc);
        done += n;
    }
    return A_OK;
}

/* ---- pager/pager_flex.h ---- */
struct pager_flex {
    gpupager_t *bank;
    uint32_t freq_hz;
    pager_flex_on_alnum_msg_func_t on_aln;
    pager_flex_on_num_msg_func_t on_num;
    pager_flex_on_siv_msg_func_t on_siv;
};

static int flex_alnum(void *user, uint32_t channel, uint16_t baud, uint8_t phase, uint8_t cycle_no, uint8_t frame_no,
                      uint64_t cap_code, int fragmented, int maildrop, uint8_t seq_num, const char *msg, size_t len)
{
    struct pager_flex *f = user;
    (void)channel;
    return f->on_aln ? f->on_aln(f, baud, phase, cycle_no, frame_no, cap_code, fragmented != 0, maildrop != 0, seq_num, msg, len) : 0;
}

static int flex_num(void *user, uint32_t channel, uint16_t baud, uint8_t phase, uint8_t cycle_no, uint8_t frame_no,
                    uint64_t cap_code, const char *msg, size_t len)
{
    struct pager_flex *f = user;
    (void)channel;
    return f->on_num ? f->on_num(f, baud, phase, cycle_no, frame_no, cap_code, msg, len) : 0;
}

static int flex_siv(void *user, uint32_t channel, uint16_t baud, uint8_t phase, uint8_t cycle_no, uint8_t frame_no,
                    uint64_t cap_code, uint8_t siv_msg_type, uint32_t data)
{
    struct pager_flex *f = user;
    (void)channel;
    return f->on_siv ? f->on_siv(f, baud, phase, cycle_no, frame_no, cap_code, siv_msg_type, data) : 0;
}

aresult_t pager_flex_new(struct pager_flex **pflex, uint32_t freq_hz, pager_flex_on_alnum_msg_func_t on_aln_msg,
        pager_flex_on_num_msg_func_t on_num_msg, pager_flex_on_siv_msg_func_t on_siv_msg)
{
    /* pager/pager_flex.c:1355-1357 */
    if (NULL == pflex || NULL == on_aln_msg || NULL == on_num_msg) return A_E_BADARGS;
    *pflex = NULL;
    struct pager_flex *f = calloc(1, sizeof(*f));
    if (NULL == f) return A_E_NOMEM;
    f->freq_hz = freq_hz; f->on_aln = on_aln_msg; f->on_num = on_num_msg; f->on_siv = on_siv_msg;
    const int rc = one_channel_bank(&f->bank, GPUPAGER_DECODER_FLEX);
    if (rc) { free(f); return map_rc(rc); }
    *pflex = f;
    return A_OK;
}

aresult_t pager_flex_delete(struct pager_flex **pflex)
{
    if (NULL == pflex || NULL == *pflex) return A_E_BADARGS;
    gpupager_destroy(&(*pflex)->bank);
    free(*pflex);
    *pflex = NULL;
    return A_OK;
}

aresult_t pager_flex_on_pcm(struct pager_flex *flex, const int16_t *pcm_samples, size_t nr_samples)
{
    if (NULL == flex || NULL == pcm_samples || 0 == nr_samples) return A_E_BADARGS;     /* pager_flex.c:1405-1407 */
    for (size_t done = 0; done < nr_samples; ) {
        const size_t n = nr_samples - done < COMPAT_MAX_FEED ? nr_samples - done : COMPAT_MAX_FEED;
        int rc = gpupager_feed(flex->bank, pcm_samples + done, n, n);
        if (!rc) rc = gpupager_dispatch_flex(flex->bank, flex_alnum, flex_num, flex_siv, flex, NULL);
        if (rc) return map_rc(rc);
        done += n;
    }
    return A_OK;
}
