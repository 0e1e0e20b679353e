"""Converts between the reference's pickled FrEIA state dicts and the .ikfw container (ikflow_b200/weight_files.py).

    python scripts/convert_weights.py MODEL.pkl MODEL.ikfw --model_name panda__full__lp191_5.25m
    python scripts/convert_weights.py MODEL.ikfw MODEL.pkl
    python scripts/convert_weights.py LIGHTNING_STATE.pkl MODEL.ikfw --model_name ... --strip_prefix   # format_state_dict first

The hyper-parameters of a .pkl come from model_descriptions.yaml (``--model_name``), as in ``get_ik_solver``; an .ikfw
file carries its own.
"""
import argparse
import os
import sys

import yaml

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ikflow_b200 import weight_files  # noqa: E402
from ikflow_b200.model import IkflowModelParameters  # noqa: E402
from ikflow_b200.robots import get_robot  # noqa: E402


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("src")
    ap.add_argument("dst")
    ap.add_argument("--model_name", help="entry of ikflow_b200/model_descriptions.yaml (needed for .pkl -> .ikfw)")
    ap.add_argument("--strip_prefix", action="store_true", help="apply format_state_dict ('nn_model.' prefix of Lightning checkpoints)")
    args = ap.parse_args()
    if args.src.endswith(".ikfw"):
        sd, params, dim_cond, ndof = weight_files.load_ikfw(args.src)
        weight_files.save_pickled_state_dict(args.dst, sd)
        print(f"{args.src}: width {params.dim_latent_space}, {params.nb_nodes} blocks, hidden {params.coeff_fn_internal_size}, dim_cond {dim_cond}, ndof {ndof} -> {args.dst}")
        return
    assert args.model_name, "--model_name is needed to read the hyper-parameters of a pickled state dict"
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with open(os.path.join(here, "ikflow_b200", "model_descriptions.yaml")) as f:
        desc = yaml.safe_load(f)[args.model_name]
    params = IkflowModelParameters()
    params.__dict__.update(desc["model_params"] if "model_params" in desc else {k: v for k, v in desc.items() if k in params.__dict__})
    robot = get_robot(desc["robot_name"])
    sd = weight_files.load_pickled_state_dict(args.src)
    if args.strip_prefix:
        sd = weight_files.format_state_dict(sd)
    dim_cond = 8 if params.softflow_enabled else 7
    weight_files.save_ikfw(args.dst, sd, params, dim_cond, robot.ndof)
    print(f"{args.src} -> {args.dst} ({os.path.getsize(args.dst)} bytes)")


if __name__ == "__main__":
    main()
