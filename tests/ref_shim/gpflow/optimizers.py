"""gpflow.optimizers.NaturalGradient (default XiNat): for each (q_mu, q_sqrt) pair
θ ← θ − γ ∂L/∂η, natural parameters θ = (S⁻¹μ, −½S⁻¹), expectation parameters η = (μ, S + μμᵀ), S = q_sqrt q_sqrtᵀ
(Salimbeni, Eleftheriadis & Hensman 2018, the algorithm GPflow implements). ∂L/∂η is obtained from the ordinary
gradients (∂L/∂μ, ∂L/∂q_sqrt) by autograd through η → (μ, chol(η₂ − η₁η₁ᵀ)), as GPflow does."""
import tensorflow as tf
import torch


class NaturalGradient:
    def __init__(self, gamma, name=None):
        self.gamma = gamma

    def minimize(self, loss_fn, var_list):
        pairs = [tuple(p[:2]) for p in var_list]
        leaves = []
        for q_mu, q_sqrt in pairs:
            leaves += [q_mu.unconstrained_variable, q_sqrt.unconstrained_variable]
        loss = loss_fn()
        tape = tf.GradientTape()
        grads = tape.gradient(loss, leaves)
        for i, (q_mu, q_sqrt) in enumerate(pairs):
            self._apply(q_mu, q_sqrt, grads[2 * i], grads[2 * i + 1])

    def _apply(self, q_mu, q_sqrt, g_mu_u, g_sqrt_u):
        mu = q_mu.value().detach().as_subclass(torch.Tensor)            # [M, D]
        R = q_sqrt.value().detach().as_subclass(torch.Tensor)           # [D, M, M]
        # ordinary gradients in constrained space (q_mu: identity; q_sqrt: FillTriangular is a permutation)
        g_mu = g_mu_u.as_subclass(torch.Tensor)
        g_R = q_sqrt.transform.forward(g_sqrt_u).as_subclass(torch.Tensor)
        eta1 = mu.clone().requires_grad_(True)                          # [M, D]
        eta2 = (R @ R.transpose(-1, -2) + torch.einsum("md,nd->dmn", mu, mu)).clone().requires_grad_(True)
        S = eta2 - torch.einsum("md,nd->dmn", eta1, eta1)
        R_of_eta = torch.linalg.cholesky(S)
        d1, d2 = torch.autograd.grad([eta1, R_of_eta], [eta1, eta2], grad_outputs=[g_mu, torch.tril(g_R)], allow_unused=True)
        # XiNat: step in the natural parameters
        Sinv = torch.linalg.inv(R @ R.transpose(-1, -2))
        nat1 = torch.einsum("dmn,nd->md", Sinv, mu) - self.gamma * d1
        nat2 = -0.5 * Sinv - self.gamma * 0.5 * (d2 + d2.transpose(-1, -2))
        S_new = torch.linalg.inv(-2.0 * nat2)
        S_new = 0.5 * (S_new + S_new.transpose(-1, -2))
        mu_new = torch.einsum("dmn,nd->md", S_new, nat1)
        q_mu.assign(mu_new)
        q_sqrt.assign(torch.linalg.cholesky(S_new))
