// inst_cartpole.cu -- cartpole architectures (sde4mbrlExamples/cartpole/cartpole_sde.yaml:12-75
// and the shipped my_models/cartpole_bb_*.pkl).
#include "registry.cuh"
namespace sde4 {
using CartD32 = Net<3, 32, 32, 1, SWISH>;
using CartD8 = Net<3, 8, 8, 1, SWISH>;
using CartResSi = Net<3, 8, 24, 2, TANH>;
using CartResBb = Net<4, 8, 24, 2, TANH>;
using CartCtrl = Net<2, 6, 8, 2, TANH>;
// noise inputs reduced_state[1,2,3], noise outputs states [1,3]
using CartSiD32 = Arch<SDE4_MODEL_CARTPOLE, 4, 1, SDE4_CART_SIDE_INFO, CartD32, 0xE, 0xA, true, CartResSi, 0, CartCtrl, 0>;
using CartBbD32 = Arch<SDE4_MODEL_CARTPOLE, 4, 1, 0, CartD32, 0xE, 0xA, true, CartResBb, 0, NoNet, 0>;
using CartBbD8 = Arch<SDE4_MODEL_CARTPOLE, 4, 1, 0, CartD8, 0xE, 0xA, true, CartResBb, 0, NoNet, 0>;
using CartSiOde = Arch<SDE4_MODEL_CARTPOLE, 4, 1, SDE4_CART_SIDE_INFO, NoNet, 0, 0, false, CartResSi, 0, CartCtrl, 0>;
using CartBbOde = Arch<SDE4_MODEL_CARTPOLE, 4, 1, 0, NoNet, 0, 0, false, CartResBb, 0, NoNet, 0>;
// control_dependent_noise (nsde.py:362): density input = [xdot, sin, cos, thetadot]/s restricted to [1,2,3], then u
using CartSiD32Cdn = Arch<SDE4_MODEL_CARTPOLE, 4, 1, SDE4_CART_SIDE_INFO, Net<4, 32, 32, 1, SWISH>, 0xE, 0xA, true, CartResSi, 0, CartCtrl, 0>;
// latency-bound single-problem MPC (BASELINE configs[1]) and small training minibatches use the lane-split kernels;
// the control net of the side-info model (2->6->8) is padded to 8 units by the evaluator (nets.cuh, EvLT::blk)
SDE4_REGISTER_TRAIN_LS(cartpole_si_d32, CartpoleModel, CartSiD32, 8)
SDE4_REGISTER_TRAIN_LS(cartpole_bb_d32, CartpoleModel, CartBbD32, 8)
SDE4_REGISTER(cartpole_bb_d8, CartpoleModel<CartBbD8>)
SDE4_REGISTER_TRAIN_LS(cartpole_si_d32_cdn, CartpoleModel, CartSiD32Cdn, 8)
SDE4_REGISTER_TRAIN(cartpole_si_ode, CartpoleModel, CartSiOde)
SDE4_REGISTER_TRAIN(cartpole_bb_ode, CartpoleModel, CartBbOde)
}  // namespace sde4
