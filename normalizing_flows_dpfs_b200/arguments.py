"""Command-line flags of the reference experiment driver (reference arguments.py:5-103): same names, types and
defaults, declared as a table.  `parse_args(args)` parses `args` when given, else sys.argv (the reference ignores
its parameter and always reads sys.argv; passing None reproduces that)."""
import argparse

# (flags, kwargs) -- hot-path flags first
_FLAGS = [
    (("--NF-dyn",), dict(action="store_true", help="dynamics normalising flow")),
    (("--NF-cond",), dict(action="store_true", help="conditional normalising-flow proposal")),
    (("--measurement",), dict(type=str, default="cos", help="|CRNVP|cos|NN|CGLOW|gaussian|")),
    (("--resampler_type",), dict(type=str, default="ot", help="|ot|soft|")),
    (("--alpha",), dict(type=float, default=0.5, help="soft-resampling mixture weight")),
    (("--epsilon",), dict(type=float, default=0.1, help="OT regularisation")),
    (("--scaling",), dict(type=float, default=0.75, help="OT epsilon-scaling factor")),
    (("--threshold",), dict(type=float, default=1e-3, help="OT convergence threshold")),
    (("--max_iter",), dict(type=int, default=100, help="OT iteration cap")),
    (("--num-particles",), dict(type=int, default=100)),
    (("--batchsize",), dict(type=int, default=32)),
    (("--hiddensize",), dict(type=int, default=32)),
    (("--sequence-length",), dict(dest="sequence_length", type=int, default=50)),
    (("--width",), dict(dest="width", type=int, default=128)),
    (("--pos-noise",), dict(dest="pos_noise", type=float, default=20.0)),
    (("--vel-noise",), dict(dest="vel_noise", type=float, default=20.0)),
    (("--init-with-true-state",), dict(action="store_true")),
    (("--trainType",), dict(dest="trainType", type=str, default="DPF", choices=["DPF", "SDPF", "UDPF"])),
    (("--block-length",), dict(dest="block_length", type=int, default=10)),
    (("--labeledRatio",), dict(type=float, default=1.0)),
    # training / bookkeeping
    (("--lr",), dict(type=float, default=1e-4)),
    (("--NF-lr",), dict(type=float, default=2.5)),
    (("--optim",), dict(type=str, default="Adam")),
    (("--num-epochs",), dict(type=int, default=500)),
    (("--seed",), dict(type=int, default=2)),
    (("--gpu",), dict(action="store_false")),
    (("--gpu-index",), dict(type=int, default=0)),
    (("--pretrain_ae",), dict(action="store_true")),
    (("--pretrain-NFcond",), dict(action="store_true")),
    (("--e2e-train",), dict(action="store_false")),
    (("--load-pretrainModel",), dict(action="store_true")),
    (("--resume",), dict(action="store_true")),
    (("--Dyn_nn",), dict(action="store_true")),
    (("--Obs_feature",), dict(action="store_false")),
    (("--split-ratio",), dict(type=float, default=0.9)),
    (("--dropout-keep-ratio",), dict(type=float, default=0.3)),
    (("--particle_std",), dict(type=float, default=0.2)),
    (("--true-pos-noise",), dict(dest="true_pos_noise", type=float, default=2.0)),
    (("--true-vel-noise",), dict(dest="true_vel_noise", type=float, default=2.0)),
    (("--testing",), dict(action="store_true")),
    (("--model-path",), dict(type=str, default="./model/e2e_model_bestval_e2e.pth")),
    (("--individual",), dict(action="store_true")),
    # conditional-Glow options (parsed for compatibility; the cGlow likelihood is outside the hot path)
    (("--x_size",), dict(type=tuple, default=(3, 8, 8))),
    (("--y_size",), dict(type=tuple, default=(3, 8, 8))),
    (("--x_hidden_channels",), dict(type=int, default=8)),
    (("--x_hidden_size",), dict(type=int, default=16)),
    (("--y_hidden_channels",), dict(type=int, default=8)),
    (("-K", "--flow_depth"), dict(type=int, default=1)),
    (("-L", "--num_levels"), dict(type=int, default=1)),
    (("--learn_top",), dict(type=bool, default=False)),
    (("--x_bins",), dict(type=float, default=256.0)),
    (("--y_bins",), dict(type=float, default=256.0)),
    # extension (not in the reference): host-free filter loop -- device-side ESS gate and in-kernel Philox draws instead of the
    # CPU generator (statistically equivalent, not the reference's random stream)
    (("--fast",), dict(action="store_true", default=False)),
    # extension: run the image encoder once on all B*T frames before the filter loop
    (("--hoist-encoder",), dict(dest="hoist_encoder", action="store_true", default=False)),
]


def build_parser():
    parser = argparse.ArgumentParser(description="NF-DPF disk tracking (B200-native hot path)")
    for flags, kw in _FLAGS:
        parser.add_argument(*flags, **kw)
    return parser


def parse_args(args=None):
    return build_parser().parse_args(args)
