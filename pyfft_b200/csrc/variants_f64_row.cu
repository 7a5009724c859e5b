// Kernel instantiations for group F64_ROW (see variants.def).
#include "kernels.cuh"

namespace b2 {
void register_f64_row(std::vector<KernelVariant>& out) {
#define B2_GROUP_F64_ROW
#define X B2_V
#define XT B2_VT
#define XC B2_VC
#define XF B2_VF
#define XCF B2_VCF
#define XB B2_VB
#define XCA B2_VCA
#define X0 B2_V0
#define XT0 B2_VT0
#define XC0 B2_VC0
#define XU B2_VU
#define XS B2_VS
#define XW B2_VW
#define XP B2_VP
#define XR B2_VR
#define XR4 B2_VR4
#define XTA B2_VTA
#include "variants.def"
#undef X
#undef XT
#undef XC
#undef XF
#undef XCF
#undef XB
#undef XCA
#undef X0
#undef XT0
#undef XC0
#undef XU
#undef XS
#undef XW
#undef XP
#undef XR
#undef XR4
#undef XTA
}
}  // namespace b2
