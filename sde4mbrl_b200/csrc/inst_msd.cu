// inst_msd.cu -- mass-spring-damper architectures
// (sde4mbrlExamples/mass_spring_damper/mass_spring_damper.yaml:9-75).
#include "registry.cuh"
namespace sde4 {
using MsdDensity32 = Net<2, 32, 32, 1, SWISH>;
// nesde_bboxes: black-box residual Fres(x) (2->4->16->1 swish), noise on qdot only
using MsdBbD32 = Arch<SDE4_MODEL_MSD, 2, 1, 0, MsdDensity32, 0x3, 0x2, true, Net<2, 4, 16, 1, SWISH>, 0, NoNet, 0>;
// side information: Fres(qdot) (1->4->16->1 swish)
using MsdSiD32 = Arch<SDE4_MODEL_MSD, 2, 1, SDE4_MSD_SIDE_INFO, MsdDensity32, 0x3, 0x2, true, Net<1, 4, 16, 1, SWISH>, 0, NoNet, 0>;
// ground truth ODE/SDE with constant diffusion, controlled
using MsdGt = Arch<SDE4_MODEL_MSD, 2, 1, SDE4_MSD_GROUND_TRUTH | SDE4_MSD_CONTROL, NoNet, 0, 0, false, NoNet, 0, NoNet, 0>;
// control_dependent_noise (nsde.py:362): density input = [q, qdot, u]
using MsdBbD32Cdn = Arch<SDE4_MODEL_MSD, 2, 1, 0, Net<3, 32, 32, 1, SWISH>, 0x3, 0x2, true, Net<2, 4, 16, 1, SWISH>, 0, NoNet, 0>;
// 4 lanes: the residual net's first layer has 4 units (BASELINE configs[0] is 200 particles: 7 warps otherwise)
SDE4_REGISTER_TRAIN_LS(msd_bb_d32, MsdModel, MsdBbD32, 4)
SDE4_REGISTER_LS(msd_si_d32, MsdModel, MsdSiD32, 4)
SDE4_REGISTER(msd_gt, MsdModel<MsdGt>)
SDE4_REGISTER_LS(msd_bb_d32_cdn, MsdModel, MsdBbD32Cdn, 4)
}  // namespace sde4
