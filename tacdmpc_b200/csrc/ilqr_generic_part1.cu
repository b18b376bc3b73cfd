// ilqr_generic.cu, part 1 of 4 (see the note at its top): fp32 backward instantiations.
#define TACD_GEN_PART 1
#include "ilqr_generic.cu"
