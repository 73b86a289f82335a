/* The FM discriminator object of multifm/fm_demod.h:22-34 on the B200 library: same three entry points, same argument
 * meaning and results (multifm/fm_demod.c:36-85, multifm/fast_atan2f.c:101-174 bit for bit), backed by gpufm_* in
 * tslb200_gpuchan.h.  There is no CPU path: _init fails when no sm_100 device is present. */
#ifndef TSLB200_COMPAT_FM_DEMOD_H
#define TSLB200_COMPAT_FM_DEMOD_H
#include <tsl/result.h>

struct demod_base;

/* fm_demod.h:22 -- new discriminator; *handle receives it */
aresult_t multifm_fm_demod_init(struct demod_base **handle);
/* fm_demod.h:28 -- nr_iq complex int16 pairs in, one int16 PCM sample each out; *nr_pcm and *nr_pcm_bytes report what was written */
aresult_t multifm_fm_demod_process(struct demod_base *handle, int16_t *iq, size_t nr_iq, int16_t *pcm, size_t *nr_pcm, size_t *nr_pcm_bytes);
/* fm_demod.h:34 -- release; *handle becomes NULL */
aresult_t multifm_fm_demod_cleanup(struct demod_base **handle);
#endif
