"""create_model(args, logger) with the reference's argument mapping for the two hot-path rows
(models/model_utils.py:44-49); every other --model string belongs to the stock reference zoo and is rejected here."""
import torch

from .armnet import ARMNetModel
from .armnet_1h import ARMNetModel as ARMNet1H


def create_model(args, logger):
    logger.info(f'=> creating model {args.model}')
    if args.model == 'armnet':
        model = ARMNetModel(args.nfield, args.nfeat, args.nemb, args.nattn_head, args.alpha, args.h,
                            args.mlp_nlayer, args.mlp_nhid, args.dropout, args.ensemble, args.dnn_nlayer,
                            args.dnn_nhid)
    elif args.model == 'armnet_1h':
        model = ARMNet1H(args.nfield, args.nfeat, args.nemb, args.alpha, args.h, args.nemb, args.mlp_nlayer,
                         args.mlp_nhid, args.dropout, args.ensemble, args.dnn_nlayer, args.dnn_nhid)
    else:
        raise ValueError(f'unknown model {args.model}')
    if torch.cuda.is_available():
        model = model.cuda()
    logger.info(f'{model}\nmodel parameters: {sum([p.data.nelement() for p in model.parameters()])}')
    return model
