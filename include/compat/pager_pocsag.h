/* The POCSAG decoder object of pager/pager_pocsag.h:8-56 on a one-channel B200 pager bank (tslb200_gpupager.h): the
 * reference's type names, entry points and argument order.  Input: int16 PCM at 38400 Hz (pager/pager_pocsag.c:132-139).
 * Callbacks run synchronously inside pager_pocsag_on_pcm, in decode order; `text` is valid during the call only
 * (pager/pager_pocsag.c:279-281). */
#ifndef TSLB200_COMPAT_PAGER_POCSAG_H
#define TSLB200_COMPAT_PAGER_POCSAG_H
#include <tsl/result.h>
#include <stdbool.h>

struct pager_pocsag;

/* pager_pocsag.h:8-22: (decoder, baud, capcode, text, length, function bits) */
typedef aresult_t (*pager_pocsag_on_numeric_msg_func_t)(struct pager_pocsag *dec, uint16_t baud, uint32_t cap, const char *text, size_t len, uint8_t func);
typedef aresult_t (*pager_pocsag_on_alpha_msg_func_t)(struct pager_pocsag *dec, uint16_t baud, uint32_t cap, const char *text, size_t len, uint8_t func);

/* pager_pocsag.h:35; the last argument is stored and never read by the reference (pager/pager_pocsag.c:185) */
aresult_t pager_pocsag_new(struct pager_pocsag **out, uint32_t channel_hz, pager_pocsag_on_numeric_msg_func_t numeric_cb,
                           pager_pocsag_on_alpha_msg_func_t alpha_cb, bool skip_bch);
/* pager_pocsag.h:45 */
aresult_t pager_pocsag_delete(struct pager_pocsag **dec);
/* pager_pocsag.h:56 */
aresult_t pager_pocsag_on_pcm(struct pager_pocsag *dec, const int16_t *pcm, size_t nr);
#endif
