"""create_model(args, logger) with the reference's argument mapping (models/model_utils.py:27-88) for the two hot-path
rows (`armnet`, `armnet_1h`, :44-49) and the config-5 zoo models that share the gather kernels (`afm`, `dcn`, `dcn+`,
`cin`, `xdfm`, `afn`, :35-43,74-79); every other --model string belongs to the stock reference zoo and is rejected."""
import torch

from . import zoo
from .armnet import ARMNetModel
from .armnet_1h import ARMNetModel as ARMNet1H

# --model string -> constructor call on the flat argparse namespace of train.py
_BUILDERS = {
    'armnet': lambda a: ARMNetModel(a.nfield, a.nfeat, a.nemb, a.nattn_head, a.alpha, a.h, a.mlp_nlayer, a.mlp_nhid,
                                    a.dropout, a.ensemble, a.dnn_nlayer, a.dnn_nhid),
    'armnet_1h': lambda a: ARMNet1H(a.nfield, a.nfeat, a.nemb, a.alpha, a.h, a.nemb, a.mlp_nlayer, a.mlp_nhid, a.dropout,
                                    a.ensemble, a.dnn_nlayer, a.dnn_nhid),       # d_k = nemb, model_utils.py:48
    'afm': lambda a: zoo.AFMModel(a.nfeat, a.nemb, a.h, a.dropout),
    'dcn': lambda a: zoo.CrossNetModel(a.nfield, a.nfeat, a.nemb, a.k),
    'dcn+': lambda a: zoo.DCNModel(a.nfield, a.nfeat, a.nemb, a.k, a.mlp_nlayer, a.mlp_nhid, a.dropout),
    'cin': lambda a: zoo.CINModel(a.nfield, a.nfeat, a.nemb, a.k, a.h),
    'xdfm': lambda a: zoo.xDeepFMModel(a.nfield, a.nfeat, a.nemb, a.k, a.h, a.mlp_nlayer, a.mlp_nhid, a.dropout),
    'afn': lambda a: zoo.AFNModel(a.nfield, a.nfeat, a.nemb, a.h, a.mlp_nlayer, a.mlp_nhid, a.dropout, a.ensemble,
                                  a.dnn_nlayer, a.dnn_nhid),
}


def create_model(args, logger):
    """Same surface as the reference factory: logs the choice, builds, moves to the GPU when there is one, logs the
    module and its parameter count; unknown names raise ValueError('unknown model ...') (model_utils.py:84)."""
    logger.info(f'=> creating model {args.model}')
    build = _BUILDERS.get(args.model)
    if build is None:
        raise ValueError(f'unknown model {args.model}')
    model = build(args)
    if torch.cuda.is_available():
        model = model.cuda()
    n_params = sum(p.numel() for p in model.parameters())
    logger.info(f'{model}\nmodel parameters: {n_params}')
    return model
