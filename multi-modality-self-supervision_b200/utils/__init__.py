from .utils import set_seed  # noqa: F401
