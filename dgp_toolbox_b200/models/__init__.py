from .dgp import DGP, DGP_Base  # noqa: F401
