/* Drop-in for pager/pager_pocsag.h:8-56: the reference's POCSAG decoder object, callbacks and argument meaning, over a
 * one-channel B200 pager bank (tslb200_gpupager.h).  Input is int16 PCM at 38400 Hz (pager/pager_pocsag.c:132-139).
 * Callbacks fire synchronously inside pager_pocsag_on_pcm, in decode order; `data` is only valid during the call
 * (pager/pager_pocsag.c:279-281). */
#pragma once

#include <tsl/result.h>
#include <stdbool.h>

struct pager_pocsag;

typedef aresult_t (*pager_pocsag_on_numeric_msg_func_t)(
        struct pager_pocsag *pocsag,
        uint16_t baud_rate,
        uint32_t capcode,
        const char *data,
        size_t data_len,
        uint8_t function);

typedef aresult_t (*pager_pocsag_on_alpha_msg_func_t)(
        struct pager_pocsag *pocsag,
        uint16_t baud_rate,
        uint32_t capcode,
        const char *data,
        size_t data_len,
        uint8_t function);

/* pager/pager_pocsag.h:35; skip_bch_decode is stored and never read by the reference (pager/pager_pocsag.c:185) */
aresult_t pager_pocsag_new(struct pager_pocsag **ppocsag, uint32_t freq_hz, pager_pocsag_on_numeric_msg_func_t on_numeric,
        pager_pocsag_on_alpha_msg_func_t on_alpha, bool skip_bch_decode);

/* pager/pager_pocsag.h:45 */
aresult_t pager_pocsag_delete(struct pager_pocsag **ppocsag);

/* pager/pager_pocsag.h:56 */
aresult_t pager_pocsag_on_pcm(struct pager_pocsag *pocsag, const int16_t *pcm_samples, size_t nr_samples);
