"""Deterministic synthetic weights for golden fixtures (test infrastructure, not product code).

The default network has 2.4 M parameters: storing a state_dict next to every golden file would put
~10 MB into the repository.  Instead the fixture generator and the tests fill the SAME architecture
from the same numpy PCG64 stream (stable across numpy/torch versions, unlike torch initialisers), so
a fixture only has to hold inputs and the reference model's outputs.
"""

import numpy as np
import torch


def fill_(module, seed):
    """In place: every parameter ~ U(-a, a) with a = sqrt(3 / fan_in) (unit-variance-preserving),
    biases ~ U(-0.1, 0.1), BatchNorm weight ~ U(0.6, 1.4), bias ~ U(-0.2, 0.2), running_mean ~
    U(-0.3, 0.3), running_var ~ U(0.5, 1.5); tensors visited in state_dict order."""
    rng = np.random.default_rng(seed)
    sd = module.state_dict()
    for name, t in sd.items():
        if name.endswith("num_batches_tracked"):
            continue
        shape = tuple(t.shape)
        leaf = name.rsplit(".", 1)[-1]
        is_bn = ".bn" in "." + name or name.startswith("bn.") or "_bn." in name
        if is_bn and leaf == "weight":
            v = rng.uniform(0.6, 1.4, size=shape)
        elif is_bn and leaf == "bias":
            v = rng.uniform(-0.2, 0.2, size=shape)
        elif leaf == "running_mean":
            v = rng.uniform(-0.3, 0.3, size=shape)
        elif leaf == "running_var":
            v = rng.uniform(0.5, 1.5, size=shape)
        elif leaf == "bias":
            v = rng.uniform(-0.1, 0.1, size=shape)
        else:
            fan_in = int(np.prod(shape[1:])) if len(shape) > 1 else shape[0]
            a = np.sqrt(3.0 / fan_in)
            v = rng.uniform(-a, a, size=shape)
        t.copy_(torch.from_numpy(v.astype(np.float32)))
    module.load_state_dict(sd)
    return module
