// tcr_net_all.cu — single translation unit for the network kernels: the resident step kernel (tcr_resident.cu) reuses the
// device bodies of the per-layer kernels, so they must be visible to one compilation (no relocatable device code needed).
#include "tcr_net_fwd.cu"
#include "tcr_net_bwd.cu"
#include "tcr_optim.cu"
#include "tcr_resident.cu"
