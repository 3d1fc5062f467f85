"""gpflow.utilities: bijector factories and set_trainable."""
import tensorflow_probability as tfp


def positive(lower=None):
    """Softplus; with a lower bound: Chain([Shift(lower), Softplus]) (GPflow 2.0 `positive(lower=...)`)."""
    if lower is None or lower == 0.0:
        return tfp.bijectors.Softplus()
    return tfp.bijectors.Chain([tfp.bijectors.Shift(lower), tfp.bijectors.Softplus()])


def triangular():
    return tfp.bijectors.FillTriangular()


def set_trainable(model, flag):
    from .base import Module, Parameter
    if isinstance(model, Parameter):
        model.trainable = flag
        return
    for p in model.parameters:
        p.trainable = flag
