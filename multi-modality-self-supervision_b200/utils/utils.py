"""`set_seed` — mirror of /root/reference/utils/utils.py:9-16 (the only helper the pre-training path uses)."""
import random

import numpy as np
import torch


_DETERMINISTIC = False


def deterministic_requested():
    """True once set_seed() ran: the reference's set_seed switches cuDNN to deterministic algorithms (utils/utils.py:14-15);
    here it additionally selects the engine's ordered (bit-reproducible) attention-dQ reduction (MV_FLAG_DETERMINISTIC)."""
    return _DETERMINISTIC


def set_seed(seed):
    global _DETERMINISTIC
    _DETERMINISTIC = True
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = False
