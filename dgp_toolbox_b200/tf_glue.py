"""TensorFlow-side glue (NOT exercised in the build image: TensorFlow is absent there).

For a TensorFlow training loop like the reference's (`tape.gradient` + `tf.optimizers.Adam.apply_gradients`,
dgp_dace/models/dgp.py:132-154) the model's trainable parameters are mirrored as `tf.Variable`s holding the UNCONSTRAINED values
(what GPflow hands to the optimiser); `loss_and_grads` runs one dgp_elbo_grad call and returns -ELBO and its gradients w.r.t.
those variables, exchanged through DLPack (zero copy when TensorFlow and torch share the GPU).

    vars_ = TFVariables(model)
    opt = tf.optimizers.Adam(0.01)
    for step in range(iterations):
        loss, grads = vars_.loss_and_grads((X, Y))
        opt.apply_gradients(zip(grads, vars_.variables))
"""
from __future__ import annotations

import torch


def _tf():
    import tensorflow as tf  # deferred: the package itself never needs TensorFlow
    return tf


def to_torch(x):
    """Eager tf.Tensor / tf.Variable -> torch tensor sharing the memory."""
    tf = _tf()
    return torch.utils.dlpack.from_dlpack(tf.experimental.dlpack.to_dlpack(tf.convert_to_tensor(x)))


def to_tf(t: torch.Tensor):
    """torch tensor -> tf.Tensor sharing the memory."""
    tf = _tf()
    return tf.experimental.dlpack.from_dlpack(torch.utils.dlpack.to_dlpack(t.contiguous()))


class TFVariables:
    """tf.Variable mirrors of a DGP_Base's trainable parameters in unconstrained space."""

    def __init__(self, model):
        tf = _tf()
        self.model = model
        self.params = list(model.trainable_parameters)
        self.variables = [tf.Variable(to_tf(p.unconstrained()), dtype=tf.float64, name=p.name or "param") for p in self.params]

    def push(self):
        """Write the variables' current values into the model (constrained space)."""
        for p, v in zip(self.params, self.variables):
            p.set_unconstrained(to_torch(v).to(p.value.device).clone())

    def loss_and_grads(self, data, seed=None):
        """(-ELBO as a tf scalar, [d(-ELBO)/d variable]) from one ELBO+gradient call of the CUDA library."""
        self.push()
        flat = self.model.elbo_flat(data, want_grad=True, seed=seed)
        grads = self.model.unpack_grads(flat)
        out = [to_tf(-p.grad_to_unconstrained(grads[p]).reshape(p.value.shape)) for p in self.params]
        return to_tf(-(flat[0] - flat[1]).reshape(())), out


def elbo_op(model, data, variables: TFVariables):
    """tf.custom_gradient wrapper: a differentiable scalar ELBO of `variables.variables` usable under tf.GradientTape
    (wrap in tf.py_function inside a tf.function; DLPack needs eager tensors)."""
    tf = _tf()

    @tf.custom_gradient
    def _elbo(*vs):
        loss, grads = variables.loss_and_grads(data)

        def grad(dy):
            return [-dy * g for g in grads]
        return -loss, grad
    return _elbo(*variables.variables)
