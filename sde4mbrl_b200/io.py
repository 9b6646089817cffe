"""Model I/O: the reference's ``{'sde': params, 'nominal': model_cfg}`` pickles
(``sde4mbrl/train_sde.py:28-33,466-469``; loaders ``sde_rotor_model.py:551-692``,
``cartpole_sde.py:115-163``) and the flattened ``.npz`` form used for committed fixtures."""
import json
import pickle

import numpy as np


def load_sde_pickle(path):
    """-> (nn_params tree {module: {name: array}}, nominal model dict)."""
    with open(path, "rb") as f:
        d = pickle.load(f)
    return d["sde"], d["nominal"]


def load_npz_fixture(path):
    """Inverse of tests/golden/make_golden.py:export_weights."""
    z = np.load(path)
    tree = {}
    nominal = None
    for k in z.files:
        if k == "__nominal_json__":
            nominal = json.loads(bytes(z[k]).decode())
            continue
        mod, name = k.split("::")
        tree.setdefault(mod, {})[name] = z[k]
    return tree, nominal


def update_params(d, u):
    """``sde4mbrl/utils.py`` update_params: deep-copying nested dict update."""
    import collections.abc
    import copy
    d = copy.deepcopy(d)
    for k, v in u.items():
        if isinstance(v, collections.abc.Mapping):
            d[k] = update_params(d.get(k, {}), v)
        else:
            d[k] = v
    return d
