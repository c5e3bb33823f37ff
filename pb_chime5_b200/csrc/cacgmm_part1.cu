// Second translation unit of the CACGMM EM kernel instantiations (see cacgmm.cu).
#define GSS_EM_PART 1
#include "cacgmm.cu"
