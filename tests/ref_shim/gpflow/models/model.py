"""gpflow.models.model: only names the reference imports (never instantiated on the hot path)."""
from typing import Tuple

import tensorflow as tf

from ..base import Module

MeanAndVariance = Tuple[tf.Tensor, tf.Tensor]


class GPModel(Module):
    pass
