// inst_hexa.cu -- hexacopter (6 rotors) SDE architectures
// (sde4mbrlExamples/rotor_uav/hexa_ahg/optimizer_sde.yaml:12-90, my_models/hexa_ahg_v4_sde.pkl).
#include "registry.cuh"
namespace sde4 {
using ResF = Net<3, 8, 16, 3, TANH>;
using ResM = Net<6, 8, 16, 3, TANH>;
// YAML network sizes: density 6->32->32->1 swish
using HexaD32s = Arch<SDE4_MODEL_ROTOR, 13, 6, SDE4_ROTOR_AERO_DRAG, Net<6, 32, 32, 1, SWISH>, ROTOR_NOISE_MASK, ROTOR_NOISE_MASK, true, ResF, 0x7, ResM, 0x3F>;
// shipped learned weights: density 6->16->16->1 tanh
using HexaD16t = Arch<SDE4_MODEL_ROTOR, 13, 6, SDE4_ROTOR_AERO_DRAG, Net<6, 16, 16, 1, TANH>, ROTOR_NOISE_MASK, ROTOR_NOISE_MASK, true, ResF, 0x7, ResM, 0x3F>;
// (2 lanes measured with scripts/lane_sweep.py: never better than 1 or 4 lanes -- not compiled)
SDE4_REGISTER_LS2_TRAIN(hexa_d32_swish, RotorModel, HexaD32s, 4, 8, 8)
SDE4_REGISTER_TRAIN_LS(hexa_d16_tanh, RotorModel, HexaD16t, 8)
}  // namespace sde4
