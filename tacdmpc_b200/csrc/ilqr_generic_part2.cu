// ilqr_generic.cu, part 2 of 4 (see the note at its top): fp64 forward instantiations.
#define TACD_GEN_PART 2
#include "ilqr_generic.cu"
