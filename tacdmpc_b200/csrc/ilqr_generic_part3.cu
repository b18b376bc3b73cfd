// ilqr_generic.cu, part 3 of 4 (see the note at its top): fp64 backward instantiations.
#define TACD_GEN_PART 3
#include "ilqr_generic.cu"
