/* Drop-in for multifm/fm_demod.h:22-34: same three entry points, same argument meaning, same results bit for bit
 * (multifm/fm_demod.c:36-85, multifm/fast_atan2f.c:101-174) -- computed by the B200 library (gpufm_* in
 * tslb200_gpuchan.h).  There is no CPU path: _init fails when no sm_100 device is present. */
#pragma once

#include <tsl/result.h>

struct demod_base;

/* multifm/fm_demod.h:22 */
aresult_t multifm_fm_demod_init(struct demod_base **pdemod);

/* multifm/fm_demod.h:28: nr_in_samples complex int16 pairs in, one int16 PCM sample each out */
aresult_t multifm_fm_demod_process(struct demod_base *demod, int16_t *in_samples, size_t nr_in_samples,
        int16_t *out_samples, size_t *pnr_out_samples, size_t *pnr_out_bytes);

/* multifm/fm_demod.h:34 */
aresult_t multifm_fm_demod_cleanup(struct demod_base **pdemod);
