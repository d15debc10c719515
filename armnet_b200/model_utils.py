"""create_model(args, logger) with the reference's argument mapping (models/model_utils.py:27-88) for the two hot-path
rows (`armnet`, `armnet_1h`, :44-49) and the config-5 zoo models that share the gather kernels (`afm`, `dcn`, `dcn+`,
`cin`, `xdfm`, `afn`, :35-43,74-79); every other --model string belongs to the stock reference zoo and is rejected."""
import torch

from .armnet import ARMNetModel
from .armnet_1h import ARMNetModel as ARMNet1H
from .zoo import AFMModel, AFNModel, CINModel, CrossNetModel, DCNModel, xDeepFMModel


def create_model(args, logger):
    logger.info(f'=> creating model {args.model}')
    if args.model == 'armnet':
        model = ARMNetModel(args.nfield, args.nfeat, args.nemb, args.nattn_head, args.alpha, args.h,
                            args.mlp_nlayer, args.mlp_nhid, args.dropout, args.ensemble, args.dnn_nlayer,
                            args.dnn_nhid)
    elif args.model == 'armnet_1h':
        model = ARMNet1H(args.nfield, args.nfeat, args.nemb, args.alpha, args.h, args.nemb, args.mlp_nlayer,
                         args.mlp_nhid, args.dropout, args.ensemble, args.dnn_nlayer, args.dnn_nhid)
    elif args.model == 'afm':
        model = AFMModel(args.nfeat, args.nemb, args.h, args.dropout)
    elif args.model == 'dcn':
        model = CrossNetModel(args.nfield, args.nfeat, args.nemb, args.k)
    elif args.model == 'dcn+':
        model = DCNModel(args.nfield, args.nfeat, args.nemb, args.k, args.mlp_nlayer, args.mlp_nhid, args.dropout)
    elif args.model == 'cin':
        model = CINModel(args.nfield, args.nfeat, args.nemb, args.k, args.h)
    elif args.model == 'xdfm':
        model = xDeepFMModel(args.nfield, args.nfeat, args.nemb, args.k, args.h, args.mlp_nlayer, args.mlp_nhid,
                             args.dropout)
    elif args.model == 'afn':
        model = AFNModel(args.nfield, args.nfeat, args.nemb, args.h, args.mlp_nlayer, args.mlp_nhid, args.dropout,
                         args.ensemble, args.dnn_nlayer, args.dnn_nhid)
    else:
        raise ValueError(f'unknown model {args.model}')
    if torch.cuda.is_available():
        model = model.cuda()
    logger.info(f'{model}\nmodel parameters: {sum([p.data.nelement() for p in model.parameters()])}')
    return model
