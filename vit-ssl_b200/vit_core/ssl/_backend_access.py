"""Short import path to the backend for the ssl sub-packages."""
from .._backend import functional as Fb  # noqa: F401
from .._backend import ops  # noqa: F401
from .._backend import dp  # noqa: F401
