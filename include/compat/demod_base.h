/* Drop-in for multifm/demod_base.h:3 -- the empty tag every demodulator embeds first. */
#pragma once

struct demod_base {

};
