"""Design study (CPU only): which tensor-core operand format keeps the 96-GEMM inverse pass within 1e-4 abs of fp32?

Emulates the hidden 1024x1024 layers with split-precision operands (products of 16-bit values are exact in
fp32/fp64, so splitting + an fp32/fp64 matmul reproduces what an fp32-accumulating MMA would compute up to
accumulation order) and compares the final joint angles with an fp64 evaluation of the same network.
The tiny first/last layer of every subnet is kept in plain fp32 (SIMT in the kernel).

Run:  python scripts/precision_study.py [--batch 256] [--stress]
Result recorded in DESIGN.md section "Number format".
"""

import argparse
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import freia_flow, jrl_kinematics as jk  # noqa: E402  (design study = test infrastructure)
from ikflow_b200.model import IkflowModelParameters, make_synthetic_state_dict  # noqa: E402


def split(x, fmt, lo_scale):
    hi = x.to(fmt).to(torch.float32)
    lo = ((x - hi) * lo_scale).to(fmt).to(torch.float32)
    return hi, lo


def linear_emul(h, w, b, mode, acc_dtype):
    if mode == "fp32":
        return torch.nn.functional.linear(h, w, b)
    if mode == "bf16x1":
        return (h.bfloat16().float().to(acc_dtype) @ w.bfloat16().float().to(acc_dtype).t()).float() + b
    if mode == "fp16x1":
        return (h.half().float().to(acc_dtype) @ w.half().float().to(acc_dtype).t()).float() + b
    if mode == "tf32x1":
        def tf32(x):
            return (x.view(torch.int32) + 0x1000 & ~0x1FFF).view(torch.float32)
        return (tf32(h.contiguous()).to(acc_dtype) @ tf32(w.contiguous()).to(acc_dtype).t()).float() + b
    if mode == "fp16x3_ps":  # power-of-two prescale (weights -> max in [256,512), activations x16), ONE accumulator
        import math
        kw = 2.0 ** (8 - math.floor(math.log2(float(w.abs().max()))))
        ka = 16.0
        hs, ws = h * ka, w * kw
        hh = hs.half().float(); hl = (hs - hh).half().float()
        wh = ws.half().float(); wl = (ws - wh).half().float()
        hh, hl, wh, wl = (t.to(acc_dtype) for t in (hh, hl, wh, wl))
        acc = hh @ wh.t() + hh @ wl.t() + hl @ wh.t()
        return acc.float() / (ka * kw) + b
    fmt, scale = {"bf16x3": (torch.bfloat16, 1.0), "fp16x3": (torch.float16, 2048.0), "fp16x3_noscale": (torch.float16, 1.0),
                  "bf16x3_s": (torch.bfloat16, 256.0), "fp16x2w": (torch.float16, 2048.0)}[mode]
    hh, hl = split(h, fmt, scale)
    wh, wl = split(w, fmt, scale)
    hh, hl, wh, wl = (t.to(acc_dtype) for t in (hh, hl, wh, wl))
    main = hh @ wh.t()
    corr = hh @ wl.t() + hl @ wh.t()
    if mode == "fp16x2w":  # weights split only (activations single fp16)
        corr = hh @ wl.t()
    return (main.float() + corr.float() / scale) + b


def subnet(sd, prefix, x, mode, acc_dtype):
    h = x
    for li in range(4):
        w, b = sd[f"{prefix}.{2*li}.weight"], sd[f"{prefix}.{2*li}.bias"]
        if li in (0, 3):
            h = torch.nn.functional.linear(h, w, b)
        else:
            h = linear_emul(h, w, b, mode, acc_dtype)
        if li != 3:
            h = torch.nn.functional.leaky_relu(h, 0.01)
    return h


def flow_emul(sd, latent, cond, nb_nodes, clamp, mode, acc_dtype):
    W = latent.shape[1]
    s1 = W // 2
    s2 = W - s1
    u = latent
    for i in range(nb_nodes - 1, -1, -1):
        p = f"module_list.{2+2*i}"
        x1, x2 = u[:, :s1], u[:, s1:]
        a1 = subnet(sd, p + ".subnet1", torch.cat([x1, cond], 1), mode, acc_dtype)
        s, t = a1[:, :s2], a1[:, s2:]
        y2 = (x2 - t) * torch.exp(-clamp * 0.636 * torch.atan(s))
        a2 = subnet(sd, p + ".subnet2", torch.cat([y2, cond], 1), mode, acc_dtype)
        s, t = a2[:, :s1], a2[:, s1:]
        y1 = (x1 - t) * torch.exp(-clamp * 0.636 * torch.atan(s))
        u = torch.cat([y1, y2], 1)[:, sd[f"module_list.{1+2*i}.perm_inv"]]
    return u.mm(sd["module_list.0.M_inv"])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--stress", type=float, default=1.0)
    ap.add_argument("--nb_nodes", type=int, default=12)
    args = ap.parse_args()
    torch.set_num_threads(8)
    hp = IkflowModelParameters()
    hp.nb_nodes, hp.dim_latent_space = args.nb_nodes, 7
    sd = make_synthetic_state_dict(hp, jk.PANDA.actuated_joints_limits, seed=0, stress=args.stress)
    q, poses = jk.sample_joint_angles_and_poses(jk.PANDA, args.batch, seed=1234)
    cond = torch.cat([poses, torch.zeros(args.batch, 1)], 1)
    latent = torch.randn(args.batch, 7, generator=torch.Generator().manual_seed(4321))
    sd64 = freia_flow.state_dict_to(sd, torch.float64)
    truth, _ = freia_flow.flow_inverse(sd64, latent.double(), cond.double(), hp.nb_nodes, 3, 2.5)
    ref32, _ = freia_flow.flow_inverse(sd, latent, cond, hp.nb_nodes, 3, 2.5)
    print(f"|q| max {truth.abs().max():.3f}; fp32 torch vs fp64 truth: max abs err {(ref32.double()-truth).abs().max():.3e}")
    for mode in ["fp32", "tf32x1", "bf16x1", "fp16x1", "fp16x2w", "bf16x3", "bf16x3_s", "fp16x3_noscale", "fp16x3", "fp16x3_ps"]:
        for acc in (torch.float64, torch.float32):
            out = flow_emul(sd, latent, cond, hp.nb_nodes, 2.5, mode, acc)
            print(
                f"{mode:15s} acc={str(acc)[6:]:8s} vs truth {(out.double()-truth).abs().max():.3e}   vs fp32-torch"
                f" {(out-ref32).abs().max():.3e}"
            )


if __name__ == "__main__":
    main()
