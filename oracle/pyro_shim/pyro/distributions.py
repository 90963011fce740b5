"""pyro.distributions wraps torch.distributions; the shim re-exports torch's."""
from torch.distributions import Normal, Categorical, OneHotCategorical, Uniform  # noqa: F401
