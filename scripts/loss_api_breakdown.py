"""Where the time of the full training loss (data term + distance-aware regulariser + weight penalty) goes when it is
called through create_model_loss_fn(...).loss_fn.value_and_grad (bench extra `ms_full_loss_value_and_grad_api`).
cProfile of 20 calls at the cartpole_sde.yaml shape (B = 512, H = 30, 20 ball samples)."""
import cProfile, pstats, io, contextlib, time, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from sde4mbrl_b200 import configs, nsde, sde_models, random as R

dev = torch.device("cuda:0")
BT, HT = 512, 30
pml = configs.cartpole_model(horizon=1, num_particles=1)
lpl = {"seed": 0, "stepsize": 0.02, "data_stepsize": 0.02, "num_particles": 1, "num_particles_test": 1, "horizon": HT,
       "obs_weights": [9, 5, 1, 1.0, 10], "pen_data": 1.0, "pen_weights": 1e-3, "default_weights": 1.0,
       "special_parameters_pen": {"scaler": 0, "density": 0},
       "density_loss": {"learn_mucoeff": {"type": "dnn", "init_value": 0.01, "activation_fn": "tanh", "hidden_layers": [8, 8]},
                        "mu_coeff": 10.0, "ball_radius": [0.1, 0.1, 0.4], "ball_nsamples": 20},
       "pen_grad_density": 1.0, "pen_density_scvex": 1.0, "pen_mu_type": "lin_inv", "pen_mu_coeff": 1000.0, "pen_scvex_mult": 0.01}
with contextlib.redirect_stdout(io.StringIO()):
    nnl, loss_fn, _, _ = nsde.create_model_loss_fn(pml, lpl, sde_constr=sde_models.CartPoleSDE, verbose=False)
g = torch.Generator(device=dev).manual_seed(5)
ym = torch.randn((BT, HT + 1, 4), device=dev, generator=g) * 0.5
ut = torch.rand((BT, HT, 1), device=dev, generator=g) * 2 - 1
for rep in range(3):
    loss_fn.value_and_grad(nnl, ym, ut, R.PRNGKey(rep))
torch.cuda.synchronize()
t0 = time.perf_counter()
for rep in range(20):
    loss_fn.value_and_grad(nnl, ym, ut, R.PRNGKey(rep))
torch.cuda.synchronize()
print("ms per call: %.3f" % ((time.perf_counter() - t0) / 20 * 1e3))
pr = cProfile.Profile()
pr.enable()
for rep in range(20):
    loss_fn.value_and_grad(nnl, ym, ut, R.PRNGKey(rep))
torch.cuda.synchronize()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(32)
print(s.getvalue()[:6000])
