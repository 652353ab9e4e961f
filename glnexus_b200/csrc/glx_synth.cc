// placeholder: synthetic cohort generator (filled in below)
#include "../../include/glx_host.h"
extern "C" int glxh_synth_batch(const char*, const glxh_synth_params*, glxh_batch**, char*, size_t) { return GLX_E_NOT_IMPLEMENTED; }
