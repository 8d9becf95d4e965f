"""Mirror of the hot-path pieces of the reference's ``utils`` package."""
from .data_pre import prepare_batch          # noqa: F401
from .experiment import Experiment           # noqa: F401
