"""Oracle helpers (TEST INFRASTRUCTURE - see oracle/__init__.py)."""
import numpy as np
import torch


def rbd(data: dict) -> dict:
    """upstream `lightglue.utils.rbd`: remove the batch dimension of every array-like value
    (used at `/root/reference/slam/core/features_utils.py:95,162,240-241`)."""
    return {k: v[0] if isinstance(v, (torch.Tensor, np.ndarray, list)) else v for k, v in data.items()}
