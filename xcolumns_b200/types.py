"""Container and dtype vocabulary of the boundary (mirrors xcolumns/types.py:8-28)."""
from typing import Tuple, Union

import numpy as np
import torch
from scipy.sparse import csr_matrix

Number = Union[int, float, np.number]
DType = Union[np.dtype, torch.dtype]
DenseMatrix = Union[np.ndarray, torch.Tensor]
Matrix = Union[np.ndarray, csr_matrix, torch.Tensor]
CSRMatrixAsTuple = Tuple[np.ndarray, np.ndarray, np.ndarray]

# same defaults as the reference: int32 label ids, float32 data, float64 accumulators
DefaultIndDType = np.int32
DefaultDataDType = np.float32
DefaultAccDataDType = np.float64
DefaultTorchDataDType = torch.float32
TORCH_AVAILABLE = True
