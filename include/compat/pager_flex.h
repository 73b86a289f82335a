/* Drop-in for pager/pager_flex.h:16-115: the reference's FLEX decoder object, callbacks and argument meaning, over a
 * one-channel B200 pager bank (tslb200_gpupager.h).  Input is int16 PCM at 16000 Hz (pager/pager_flex.c:1401). */
#pragma once

#include <tsl/result.h>

#include <stdint.h>
#include <stdbool.h>

struct pager_flex;

typedef aresult_t (*pager_flex_on_alnum_msg_func_t)(
        struct pager_flex *flex,
        uint16_t baud,
        uint8_t phase,
        uint8_t cycle_no,
        uint8_t frame_no,
        uint64_t cap_code,
        bool fragmented,
        bool maildrop,
        uint8_t seq_num,
        const char *message_bytes,
        size_t message_len);

typedef aresult_t (*pager_flex_on_num_msg_func_t)(
        struct pager_flex *flex,
        uint16_t baud,
        uint8_t phase,
        uint8_t cycle_no,
        uint8_t frame_no,
        uint64_t cap_code,
        const char *message_bytes,
        size_t message_len);

#define PAGER_FLEX_SIV_TEMP_ADDRESS_ACTIVATION              0x0
#define PAGER_FLEX_SIV_SYSTEM_EVENT                         0x1
#define PAGER_FLEX_SIV_RESERVED_TEST                        0x3

typedef aresult_t (*pager_flex_on_siv_msg_func_t)(
        struct pager_flex *flex,
        uint16_t baud,
        uint8_t phase,
        uint8_t cycle_no,
        uint8_t frame_no,
        uint64_t cap_code,
        uint8_t siv_msg_type,
        uint32_t data);

/* pager/pager_flex.h:95 */
aresult_t pager_flex_new(struct pager_flex **pflex, uint32_t freq_hz, pager_flex_on_alnum_msg_func_t on_aln_msg,
        pager_flex_on_num_msg_func_t on_num_msg, pager_flex_on_siv_msg_func_t on_siv_msg);

/* pager/pager_flex.h:105 */
aresult_t pager_flex_delete(struct pager_flex **pflex);

/* pager/pager_flex.h:115 */
aresult_t pager_flex_on_pcm(struct pager_flex *flex, const int16_t *pcm_samples, size_t nr_samples);
