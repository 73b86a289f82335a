/* multifm/demod_base.h:3 -- the tag type every demodulator object starts with (it carries no members). */
#ifndef TSLB200_COMPAT_DEMOD_BASE_H
#define TSLB200_COMPAT_DEMOD_BASE_H
struct demod_base { };
#endif
