"""gpflow.inducing_variables.InducingPoints. The reference's MF/MO models need a locally PATCHED GPflow whose
`InducingPoints(layers=..., Z=...)` also carries `Z_left` / `Z_right`; that patch is not in the reference repository.
Assumed semantics (from the call sites `utils/layers.py:208-213`, `MF_DGP.py:204-207,375-377`): `Z_left` is the trainable
parameter built from the given Z, `Z_right` the mean of 100 propagated samples of Z_left through the earlier layers
(the same helper the layer constructor calls), `Z = [Z_left, Z_right]`; model code later overwrites `Z_right` / `Z` with
plain tensors."""
import numpy as np
import tensorflow as tf

from .base import Module, Parameter


class InducingPoints(Module):
    def __init__(self, Z=None, layers=None, layers_red=None, name=None):
        Module.__init__(self, name=name)
        if layers is None:
            self.Z = Parameter(Z)
        else:
            self.Z_left = Z if isinstance(Z, Parameter) else Parameter(Z)
            if layers_red is None:
                from dgp_dace.utils.layers import sample_Z_right_array_all_layers
                zr = sample_Z_right_array_all_layers(layers, self.Z_left.numpy(), 100)
            else:      # embedded-mapping variant (utils/layers_red.py:110-129,162-163)
                from dgp_dace.utils.layers_red import sample_Z_right_array_all_layers
                zr = sample_Z_right_array_all_layers(layers, layers_red, self.Z_left.numpy(), 100)
            self.Z_right = tf.convert_to_tensor(zr).detach()
            self.Z = tf.concat([self.Z_left, self.Z_right], 1)

    def __len__(self):
        return int(self.Z.shape[0])

    @property
    def num_inducing(self):
        return len(self)
