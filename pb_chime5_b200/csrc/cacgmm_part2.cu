// Third translation unit of the CACGMM EM kernel instantiations (see cacgmm.cu).
#define GSS_EM_PART 2
#include "cacgmm.cu"
