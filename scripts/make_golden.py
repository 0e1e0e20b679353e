"""Generates the golden fixtures under tests/golden/ (committed).  TEST INFRASTRUCTURE.

The reference package cannot be imported here (its arithmetic lives in FrEIA 0.2 and jrl, both un-vendored and not
installable offline -- DESIGN.md), so the fixtures hold
  * the known-answer vectors of the reference's OWN tests (copied constants, with file:line), and
  * outputs of the oracle restatement (fp32 op-for-op, plus an fp64 evaluation of the same network as ground truth) on
    seeded inputs, so that the CUDA path, the oracle and future refactors are all pinned to the same numbers.
Weights are regenerated from their seed by ikflow_b200.model.make_synthetic_state_dict (a 203 MB tensor set is not a
fixture); the fixture stores a checksum of them.

Run:  python scripts/make_golden.py
"""
import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ikflow_b200.model import IkflowModelParameters, make_synthetic_state_dict  # noqa: E402
from oracle import freia_flow, jrl_kinematics as jk  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def sd_checksum(sd) -> str:
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].contiguous().numpy().tobytes())
    return h.hexdigest()


def flow_case(name, nb_nodes, width, cfg, hidden, batch, robot):
    hp = IkflowModelParameters()
    hp.nb_nodes, hp.dim_latent_space, hp.coeff_fn_config, hp.coeff_fn_internal_size = nb_nodes, width, cfg, hidden
    sd = make_synthetic_state_dict(hp, robot.actuated_joints_limits, seed=0)
    latent = torch.randn(batch, width, generator=torch.Generator().manual_seed(4321))
    _, poses = jk.sample_joint_angles_and_poses(jk.PANDA, batch, seed=1234)
    cond = torch.cat([poses, torch.zeros(batch, 1)], dim=1)
    out32, _, inter = freia_flow.flow_inverse(sd, latent, cond, nb_nodes, cfg, hp.rnvp_clamp, return_intermediates=True)
    sd64 = freia_flow.state_dict_to(sd, torch.float64)
    out64, logdet64 = freia_flow.flow_inverse(sd64, latent.double(), cond.double(), nb_nodes, cfg, hp.rnvp_clamp)
    # forward direction (x -> z with log-det, ikflow/training/lt_model.py:156) on joint-angle-like samples
    fwd_x = (torch.rand(batch, width, generator=torch.Generator().manual_seed(99)) * 2 - 1) * 2.0
    fz32, fld32 = freia_flow.flow_forward(sd, fwd_x, cond, nb_nodes, cfg, hp.rnvp_clamp)
    fz64, fld64 = freia_flow.flow_forward(sd64, fwd_x.double(), cond.double(), nb_nodes, cfg, hp.rnvp_clamp)
    np.savez_compressed(
        os.path.join(OUT, f"flow_{name}.npz"),
        fwd_x=fwd_x.numpy(), fwd_z_fp32=fz32.numpy(), fwd_z_fp64=fz64.numpy(), fwd_logdet_fp32=fld32.numpy(), fwd_logdet_fp64=fld64.numpy(),
        nb_nodes=nb_nodes, width=width, coeff_fn_config=cfg, hidden=hidden, rnvp_clamp=hp.rnvp_clamp,
        robot=robot.name, weights_sha256=sd_checksum(sd), latent=latent.numpy(), poses=poses.numpy(),
        out_fp32=out32.numpy(), out_fp64=out64.numpy(), logdet_fp64=logdet64.numpy(),
        state_after_first_block_fp32=inter[0].numpy(),
    )
    print(name, "max |fp32 - fp64|", float((out32.double() - out64).abs().max()))


def kinematics_case():
    q, poses = jk.sample_joint_angles_and_poses(jk.PANDA, 64, seed=7, dtype=torch.float64)
    q32 = q.float()
    noise = 0.05 * torch.randn(64, 7, generator=torch.Generator().manual_seed(8), dtype=torch.float64)
    seeds = jk.clamp_to_joint_limits(jk.PANDA, (q + noise).clone())
    step64 = jk.lm_step(jk.PANDA, poses, seeds.clone())
    step32 = jk.lm_step(jk.PANDA, poses.float(), seeds.float().clone())
    pe64, re64 = jk.pose_error(jk.PANDA, step64, poses)
    np.savez_compressed(
        os.path.join(OUT, "kinematics_panda.npz"),
        q=q32.numpy(), poses_fp64=poses.numpy(), fk_fp32=jk.forward_kinematics(jk.PANDA, q32).numpy(),
        lm_seeds=seeds.float().numpy(), lm_step_fp64=step64.numpy(), lm_step_fp32=step32.numpy(),
        pos_err_after_fp64=pe64.numpy(), rot_err_after_fp64=re64.numpy(),
        # reference tests/evaluation_utils_test.py:20-32
        kat_fk_zero=np.array([0.088, 0.0, 0.926, 0.0, 0.92387953, 0.38268343, 0.0]),
        kat_pos_err=1.355440887681938, kat_rot_err=3.1415927,
    )


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    flow_case("tiny_w9", 3, 9, 2, 256, 48, jk.PANDA)             # TINY_MODEL_PARAMS, ikflow/model.py:45-48
    flow_case("panda_nb12", 12, 7, 3, 1024, 96, jk.PANDA)        # panda__full__lp191_5.25m geometry
    flow_case("fetch_arm_nb16", 16, 10, 3, 1024, 40, jk.FETCH_ARM)  # fetch_arm__large__mh186_9.25m geometry
    kinematics_case()
    print(sorted(os.listdir(OUT)))
