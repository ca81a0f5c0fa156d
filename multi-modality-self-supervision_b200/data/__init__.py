from .dataset_origin import CXRDataset, truncate_txt  # noqa: F401
from .helper import get_transforms  # noqa: F401
