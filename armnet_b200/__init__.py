"""armnet_b200 -- B200-native (sm_100a) implementation of ARM-Net's forward hot path behind the reference's own
module surface. Importing the package loads libarmnet_b200.so and fails loudly if it is not built."""
from . import _capi, ops  # noqa: F401  (loads the shared library)
from .armnet import ARMNetModel, SparseAttLayer
from .armnet_1h import ARMNetModel as ARMNet1H
from .armnet_1h import SparseAttention
from .entmax import EntmaxBisect, entmax_bisect
from .layers import MLP, Embedding
from .model_utils import create_model
from .serving import BatchScorer
from . import zoo

__all__ = ['ARMNetModel', 'ARMNet1H', 'SparseAttLayer', 'SparseAttention', 'EntmaxBisect', 'entmax_bisect',
           'Embedding', 'MLP', 'create_model', 'BatchScorer', 'ops', 'zoo']
__version__ = '0.1.0'
