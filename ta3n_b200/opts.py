"""Command-line surface of the reference's opts.py (flag names, types, defaults, choices kept
identical so ``main.py``-style drivers parse the same command lines).  Table-driven; the flags
that select code outside the accelerated path are still accepted here and rejected by
``VideoModel`` with NotImplementedError.
"""
import argparse

_POSITIONAL = [
    ("class_file", dict(type=str, default="classInd.txt")),
    ("modality", dict(type=str, choices=['RGB', 'Flow', 'RGBDiff', 'RGBDiff2', 'RGBDiffplus'])),
    ("train_source_list", dict(type=str)),
    ("train_target_list", dict(type=str)),
    ("val_list", dict(type=str)),
]

YN = ['Y', 'N']
ATTN = ['none', 'TransAttn', 'general', 'DotProduct']

# (flags, kwargs) in the reference's order: opts.py:9-118
_OPTIONS = [
    # model
    (("--arch",), dict(type=str, default="resnet101")),
    (("--pretrained",), dict(type=str, default="none")),
    (("--num_segments",), dict(type=int, default=5)),
    (("--val_segments",), dict(type=int, default=-1)),
    (("--add_fc",), dict(default=1, type=int, metavar='M')),
    (("--fc_dim",), dict(type=int, default=1024)),
    (("--baseline_type",), dict(type=str, default='frame', choices=['frame', 'video', 'tsn'])),
    (("--frame_aggregation",), dict(type=str, default='avgpool',
                                    choices=['avgpool', 'rnn', 'temconv', 'trn', 'trn-m', 'none'])),
    (("--optimizer",), dict(type=str, default='SGD', choices=['SGD', 'Adam'])),
    (("--use_opencv",), dict(default=False, action="store_true")),
    (("--dropout_i", "--doi"), dict(default=0.8, type=float, metavar='DOI')),
    (("--dropout_v", "--dov"), dict(default=0.8, type=float, metavar='DOV')),
    (("--loss_type",), dict(type=str, default="nll", choices=['nll'])),
    (("--weighted_class_loss",), dict(type=str, default='N', choices=YN)),
    # rnn
    (("--n_rnn",), dict(default=1, type=int, metavar='M')),
    (("--rnn_cell",), dict(type=str, default='LSTM', choices=['LSTM', 'GRU'])),
    (("--n_directions",), dict(type=int, default=1, choices=[1, 2])),
    (("--n_ts",), dict(type=int, default=5)),
    # domain adaptation
    (("--share_params",), dict(type=str, default='Y', choices=YN)),
    (("--use_target",), dict(type=str, default='none', choices=['none', 'Sv', 'uSv'])),
    (("--dis_DA",), dict(type=str, default='none', choices=['none', 'DAN', 'JAN', 'CORAL'])),
    (("--adv_DA",), dict(type=str, default='none', choices=['none', 'RevGrad'])),
    (("--use_bn",), dict(type=str, default='none', choices=['none', 'AdaBN', 'AutoDIAL'])),
    (("--ens_DA",), dict(type=str, default='none', choices=['none', 'MCD'])),
    (("--use_attn_frame",), dict(type=str, default='none', choices=ATTN)),
    (("--use_attn",), dict(type=str, default='none', choices=ATTN)),
    (("--n_attn",), dict(type=int, default=1)),
    (("--add_loss_DA",), dict(type=str, default='none', choices=['none', 'target_entropy', 'attentive_entropy'])),
    (("--pred_normalize",), dict(type=str, default='N', choices=YN)),
    (("--alpha",), dict(default=1, type=float, metavar='M')),
    (("--beta",), dict(default=[1, 1, 1], type=float, nargs="+", metavar='M')),   # [relation, video, frame]
    (("--gamma",), dict(default=1, type=float, metavar='M')),
    (("--mu",), dict(default=0, type=float, metavar='M')),
    (("--weighted_class_loss_DA",), dict(type=str, default='N', choices=YN)),
    (("--place_dis",), dict(default=['Y', 'Y', 'N'], type=str, nargs="+", metavar='N')),
    (("--place_adv",), dict(default=['Y', 'Y', 'Y'], type=str, nargs="+", metavar='N')),
    # learning
    (("--pretrain_source",), dict(default=False, action="store_true")),
    (("--epochs",), dict(default=100, type=int, metavar='N')),
    (("-b", "--batch_size"), dict(default=[32, 28, 64], type=int, nargs="+", metavar='N')),
    (("--lr", "--learning_rate"), dict(default=0.0001, type=float, metavar='LR')),
    (("--lr_decay",), dict(default=10, type=float, metavar='LRDecay')),
    (("--lr_adaptive",), dict(type=str, default='none', choices=['none', 'loss', 'dann'])),
    (("--lr_steps",), dict(default=[60, 100], type=float, nargs="+", metavar='LRSteps')),
    (("--momentum",), dict(default=0.9, type=float, metavar='M')),
    (("--weight_decay", "--wd"), dict(default=1e-4, type=float, metavar='W')),
    (("--clip_gradient", "--gd"), dict(default=20, type=float, metavar='W')),
    (("--no_partialbn", "--npb"), dict(default=True, action="store_true")),   # stays True
    (("--copy_list",), dict(default=['N', 'Y'], type=str, nargs="+", metavar='N')),
    # monitor
    (("--print_freq", "-pf"), dict(default=10, type=int, metavar='N')),
    (("--show_freq", "-sf"), dict(default=10, type=int, metavar='N')),
    (("--eval_freq", "-ef"), dict(default=1, type=int, metavar='N')),
    (("--verbose",), dict(default=False, action="store_true")),
    # runtime
    (("-j", "--workers"), dict(default=8, type=int, metavar='N')),
    (("--resume",), dict(default='', type=str, metavar='PATH')),
    (("--resume_hp",), dict(default=False, action="store_true")),
    (("-e", "--evaluate"), dict(dest='evaluate', action='store_true')),
    (("--exp_path",), dict(type=str, default="")),
    (("--gpus",), dict(nargs='+', type=int, default=None)),
    (("--flow_prefix",), dict(default="", type=str)),
    (("--save_model",), dict(default=False, action="store_true")),
    (("--save_best_log",), dict(default="best.log", type=str)),
    (("--save_attention",), dict(type=int, default=-1)),
    (("--tensorboard",), dict(dest='tensorboard', action='store_true')),
]


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description="TA3N (B200 hot path) -- same flags as the reference's opts.py")
    for name, kw in _POSITIONAL:
        p.add_argument(name, **kw)
    for flags, kw in _OPTIONS:
        p.add_argument(*flags, **kw)
    return p


parser = build_parser()
