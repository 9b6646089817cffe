// inst_rotor_misc.cu -- remaining multirotor architectures found in the shipped pickles:
// hexacopter neural ODEs (no density net) and the quadrotor (iris) models.
#include "registry.cuh"
namespace sde4 {
using ResF = Net<3, 8, 16, 3, TANH>;
using ResM = Net<6, 8, 16, 3, TANH>;
using ResM3 = Net<3, 8, 16, 3, TANH>;
// hexa_ahg_sde / v1 / v2: ODE (or constant-diffusion SDE)
using HexaOde = Arch<SDE4_MODEL_ROTOR, 13, 6, SDE4_ROTOR_AERO_DRAG, NoNet, 0, 0, false, ResF, 0x7, ResM, 0x3F>;
// hexa_ahg_v3: residual moments see (v_b) only
using HexaOdeM3 = Arch<SDE4_MODEL_ROTOR, 13, 6, SDE4_ROTOR_AERO_DRAG, NoNet, 0, 0, false, ResF, 0x7, ResM3, 0x7>;
// hexa_ahg_linear_nonoise_v3: 24-wide residual nets
using HexaOde24 = Arch<SDE4_MODEL_ROTOR, 13, 6, SDE4_ROTOR_AERO_DRAG, NoNet, 0, 0, false, Net<3, 24, 24, 3, TANH>, 0x7, Net<6, 24, 24, 3, TANH>, 0x3F>;
// system-identification prior: physics only (load_predictor_function(prior_dist=True), sde_rotor_model.py:595-602)
using HexaPrior = Arch<SDE4_MODEL_ROTOR, 13, 6, 0, NoNet, 0, 0, false, NoNet, 0, NoNet, 0>;
// iris quadrotor
using IrisD32s = Arch<SDE4_MODEL_ROTOR, 13, 4, SDE4_ROTOR_AERO_DRAG, Net<6, 32, 32, 1, SWISH>, ROTOR_NOISE_MASK, ROTOR_NOISE_MASK, true, ResF, 0x7, ResM, 0x3F>;
using IrisOde = Arch<SDE4_MODEL_ROTOR, 13, 4, SDE4_ROTOR_AERO_DRAG, NoNet, 0, 0, false, ResF, 0x7, ResM, 0x3F>;
using IrisPrior = Arch<SDE4_MODEL_ROTOR, 13, 4, 0, NoNet, 0, 0, false, NoNet, 0, NoNet, 0>;
// control_dependent_noise (nsde.py:362): the hexacopter's density net reads [v, omega]/scale followed by the 6 rotor commands
using HexaD32sCdn = Arch<SDE4_MODEL_ROTOR, 13, 6, SDE4_ROTOR_AERO_DRAG, Net<12, 32, 32, 1, SWISH>, ROTOR_NOISE_MASK, ROTOR_NOISE_MASK, true, ResF, 0x7, ResM, 0x3F>;
SDE4_REGISTER(hexa_d32_swish_cdn, RotorModel<HexaD32sCdn>)
SDE4_REGISTER_TRAIN(hexa_ode, RotorModel, HexaOde)
SDE4_REGISTER(hexa_ode_m3, RotorModel<HexaOdeM3>)
SDE4_REGISTER(hexa_ode_w24, RotorModel<HexaOde24>)
SDE4_REGISTER(hexa_prior, RotorModel<HexaPrior>)
SDE4_REGISTER_TRAIN_LS(iris_d32_swish, RotorModel, IrisD32s, 8)
SDE4_REGISTER_TRAIN(iris_ode, RotorModel, IrisOde)
SDE4_REGISTER(iris_prior, RotorModel<IrisPrior>)
}  // namespace sde4
