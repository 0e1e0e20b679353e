"""GPU precision study: max abs / relative error of the flow kernel's operand formats against the oracle (fp32 on the CPU,
fp32 through torch-CUDA, fp64) for default-init weights and amplified last layers.  Prints one line per case.

    python scripts/precision_gpu.py [batch]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ikflow_b200  # noqa: E402
from ikflow_b200.model import IkflowModelParameters, make_synthetic_state_dict  # noqa: E402
from oracle import freia_flow, jrl_kinematics as jk  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 256
torch.backends.cuda.matmul.allow_tf32 = False
hp = IkflowModelParameters()
hp.nb_nodes, hp.dim_latent_space = 12, 7
robot = ikflow_b200.Panda()
latent = torch.randn(batch, 7, generator=torch.Generator().manual_seed(4321))
_, poses = jk.sample_joint_angles_and_poses(jk.PANDA, batch, seed=1234)
cond = torch.cat([poses, torch.zeros(batch, 1)], dim=1)
for stress in (1.0, 1.5, 2.0, 3.0):
    sd = make_synthetic_state_dict(hp, robot.actuated_joints_limits, seed=0, stress=stress)
    ref32 = freia_flow.flow_inverse(sd, latent, cond, 12, 3, 2.5)[0]
    ref64 = freia_flow.flow_inverse(freia_flow.state_dict_to(sd, torch.float64), latent.double(), cond.double(), 12, 3, 2.5)[0]
    sd_cu = freia_flow.state_dict_to(sd, device="cuda")
    ref32cu = freia_flow.flow_inverse(sd_cu, latent.cuda(), cond.cuda(), 12, 3, 2.5)[0].cpu()
    den = 1 + ref64.abs()
    print(f"stress {stress}: |q| max {ref64.abs().max():.4g}")
    print(f"   torch fp32 CPU  vs fp64: abs {(ref32.double() - ref64).abs().max():.3e} rel {((ref32.double() - ref64).abs() / den).max():.3e}")
    print(f"   torch fp32 CUDA vs fp64: abs {(ref32cu.double() - ref64).abs().max():.3e} rel {((ref32cu.double() - ref64).abs() / den).max():.3e}   vs fp32 CPU abs {(ref32cu - ref32).abs().max():.3e}")
    for precision in ("bf16x3", "fp16x3"):
        model = ikflow_b200.glow_cNF_model(hp, robot, 8, 7, precision=precision)
        model.load_state_dict(sd)
        out = model.inverse(latent.cuda(), cond.cuda()).cpu()
        e64, e32 = (out.double() - ref64).abs(), (out - ref32).abs()
        print(f"   {precision:7s} vs fp64: abs {e64.max():.3e} rel {(e64 / den).max():.3e}   vs fp32 CPU: abs {e32.max():.3e} rel {(e32 / (1 + ref32.abs())).max():.3e}   status {model.status()}  {model.last_kernel().split('kernel')[-1]}")
        del model
