/* The FLEX decoder object of pager/pager_flex.h:16-115 on a one-channel B200 pager bank (tslb200_gpupager.h): the
 * reference's type names, entry points and argument order.  Input: int16 PCM at 16000 Hz (pager/pager_flex.c:1401). */
#ifndef TSLB200_COMPAT_PAGER_FLEX_H
#define TSLB200_COMPAT_PAGER_FLEX_H
#include <tsl/result.h>
#include <stdbool.h>
#include <stdint.h>

struct pager_flex;

/* pager_flex.h:16-33: alphanumeric page */
typedef aresult_t (*pager_flex_on_alnum_msg_func_t)(struct pager_flex *dec, uint16_t baud, uint8_t phase, uint8_t cycle, uint8_t frame, uint64_t cap,
                                                    bool fragmented, bool maildrop, uint8_t seq, const char *text, size_t len);
/* pager_flex.h:35-49: numeric page */
typedef aresult_t (*pager_flex_on_num_msg_func_t)(struct pager_flex *dec, uint16_t baud, uint8_t phase, uint8_t cycle, uint8_t frame, uint64_t cap,
                                                  const char *text, size_t len);
/* pager_flex.h:51-83: short instruction vectors and their types */
#define PAGER_FLEX_SIV_TEMP_ADDRESS_ACTIVATION 0x0
#define PAGER_FLEX_SIV_SYSTEM_EVENT            0x1
#define PAGER_FLEX_SIV_RESERVED_TEST           0x3
typedef aresult_t (*pager_flex_on_siv_msg_func_t)(struct pager_flex *dec, uint16_t baud, uint8_t phase, uint8_t cycle, uint8_t frame, uint64_t cap,
                                                  uint8_t siv_type, uint32_t payload);

/* pager_flex.h:95 -- the alphanumeric and numeric callbacks are required (pager/pager_flex.c:1355-1357) */
aresult_t pager_flex_new(struct pager_flex **out, uint32_t channel_hz, pager_flex_on_alnum_msg_func_t alnum_cb,
                         pager_flex_on_num_msg_func_t num_cb, pager_flex_on_siv_msg_func_t siv_cb);
/* pager_flex.h:105 */
aresult_t pager_flex_delete(struct pager_flex **dec);
/* pager_flex.h:115 */
aresult_t pager_flex_on_pcm(struct pager_flex *dec, const int16_t *pcm, size_t nr);
#endif
