"""CPU oracle for the host logic of ``ikflow/ikflow_solver.py`` (TEST INFRASTRUCTURE).

``OracleSolver`` restates ``IKFlowSolver`` on top of ``oracle.freia_flow`` and ``oracle.jrl_kinematics``
op for op -- including the per-row Python selection loop (``ikflow_solver.py:217-222``, "last valid
repeat wins"), the boolean-mask compaction (``:231-232``) and the retry schedule (``:387-408``) -- so
that (a) the CUDA path can be compared against it on identical latent draws and (b) timing it is a
faithful stand-in for the reference's own torch path (on CPU: ``bench.py --impl reference``).

The only liberty taken: the latent draw of ``_generate_exact_ik_solutions`` (hard-coded gaussian,
``:165-166,187``) goes through ``self.latent_source`` so tests can inject the same draws into both
implementations.  By default it is ``draw_latent`` exactly as in the reference.
"""

from typing import Callable, Dict, Optional, Tuple

import torch

from . import freia_flow, jrl_kinematics as jk


def draw_latent(latent_distribution: str, latent_scale: float, shape, device) -> torch.Tensor:
    """``ikflow/ikflow_solver.py:16-29``."""
    assert latent_distribution in ["gaussian", "uniform"]
    assert latent_scale > 0
    assert len(shape) == 2
    if latent_distribution == "gaussian":
        return latent_scale * torch.randn(shape, device=device)
    return 2 * latent_scale * torch.rand(shape, device=device) - latent_scale


class OracleSolver:
    def __init__(
        self,
        robot: jk.ChainRobot,
        state_dict: Dict[str, torch.Tensor],
        nb_nodes: int,
        dim_latent_space: int,
        coeff_fn_config: int = 3,
        rnvp_clamp: float = 2.5,
        softflow_enabled: bool = True,
        device: str = "cpu",
        latent_source: Optional[Callable] = None,
    ):
        self.robot = robot
        self.ndof = robot.ndof
        self.dim_cond = 8 if softflow_enabled else 7  # ikflow_solver.py:51-53
        self.network_width = dim_latent_space
        self.nb_nodes = nb_nodes
        self.coeff_fn_config = coeff_fn_config
        self.rnvp_clamp = rnvp_clamp
        self.device = device
        self.sd = freia_flow.state_dict_to(state_dict, device=device)
        self.latent_source = latent_source or (lambda shape, dev: draw_latent("gaussian", 1.0, shape, dev))

    # ikflow_solver.py:85-110
    def _run_inference(self, latent, conditional, clamp_to_joint_limits: bool):
        assert latent.shape[0] == conditional.shape[0]
        output_rev, _ = freia_flow.flow_inverse(
            self.sd, latent, conditional, self.nb_nodes, self.coeff_fn_config, self.rnvp_clamp
        )
        solutions = output_rev[:, 0 : self.ndof]
        if clamp_to_joint_limits:
            solutions = jk.clamp_to_joint_limits(self.robot, solutions)
        return solutions

    # ikflow_solver.py:254-343
    def generate_ik_solutions(
        self,
        y: torch.Tensor,
        n: Optional[int] = None,
        latent: Optional[torch.Tensor] = None,
        latent_distribution: str = "gaussian",
        latent_scale: float = 1.0,
        clamp_to_joint_limits: bool = True,
    ) -> torch.Tensor:
        if y.numel() == 7:
            assert isinstance(n, int) and n > 0
        else:
            assert y.shape[1] == 7
        n = y.shape[0] if n is None else n
        device = y.device
        with torch.inference_mode():
            zeros = torch.zeros((n, self.dim_cond - 7), dtype=torch.float32, device=device)
            if y.numel() == 7:
                conditional = torch.cat([y.expand((n, 7)), zeros], dim=1)
            else:
                conditional = torch.cat([y, zeros], dim=1)
            if latent is None:
                latent = draw_latent(latent_distribution, latent_scale, (n, self.network_width), device)
            return self._run_inference(latent, conditional, clamp_to_joint_limits)

    # ikflow_solver.py:119-247
    def _generate_exact_ik_solutions(
        self,
        target_poses: torch.Tensor,
        repeat_count: int,
        n_opt_steps_max: int,
        pos_error_threshold: float,
        rot_error_threshold: float,
        run_lma_on_cpu: bool = False,
    ) -> Tuple[torch.Tensor, torch.Tensor]:
        n = target_poses.shape[0]
        n_tiled = n * repeat_count
        device = target_poses.device
        device_0 = device
        do_run_entirely_on_cpu = run_lma_on_cpu and n < 750

        with torch.inference_mode():
            conditional = torch.cat(
                [target_poses, torch.zeros((n, self.dim_cond - 7), dtype=torch.float32, device=device)], dim=1
            )
            conditional_tiled = conditional.repeat((repeat_count, 1))
            target_poses_tiled = conditional_tiled[:, 0:7]
            latent = self.latent_source((n_tiled, self.network_width), device)
            q = self._run_inference(latent, conditional_tiled, True)

            if run_lma_on_cpu and do_run_entirely_on_cpu:
                q = q.cpu()
                target_poses_tiled = target_poses_tiled.cpu()
                device = "cpu"

            final_solutions = torch.zeros(n, self.ndof, dtype=torch.float32, device=device)
            final_valids = torch.zeros(n, dtype=torch.bool, device=device)
            n_invalid = n

            for _ in range(n_opt_steps_max):
                assert len(q) == n_invalid * repeat_count
                if run_lma_on_cpu and not do_run_entirely_on_cpu:
                    q = jk.lm_step(self.robot, target_poses_tiled.cpu(), q.cpu())
                    q = q.to(device)
                else:
                    q = jk.lm_step(self.robot, target_poses_tiled, q)
                pos_errors, rot_errors = jk.pose_error(self.robot, q, target_poses_tiled)
                valids_i_tiled = torch.logical_and(pos_errors < pos_error_threshold, rot_errors < rot_error_threshold)

                valids_i = torch.zeros(n_invalid, dtype=torch.bool, device=device)
                sols_i = torch.zeros((n_invalid, self.ndof), dtype=torch.float32, device=device)
                valid_idxs = torch.nonzero(valids_i_tiled)
                for j in range(valid_idxs.shape[0]):  # the reference's per-row loop; later idx overwrites earlier
                    idx = valid_idxs[j, 0]
                    sol_idx = idx % n_invalid
                    sols_i[sol_idx, :] = q[idx, :]
                    valids_i[sol_idx] = True

                final_solutions[torch.logical_not(final_valids)] = sols_i
                final_valids[torch.logical_not(final_valids)] = valids_i

                if final_valids.all():
                    return final_solutions.to(device_0), final_valids.to(device_0)

                keep = torch.logical_not(valids_i).repeat((repeat_count))
                q = q[keep, :]
                target_poses_tiled = target_poses_tiled[keep, :]
                n_invalid = n - final_valids.sum().item()

            return final_solutions.to(device_0), final_valids.to(device_0)

    # ikflow_solver.py:345-411
    def generate_exact_ik_solutions(
        self,
        target_poses: torch.Tensor,
        repeat_counts: Tuple[int, ...] = (1, 3, 10),
        pos_error_threshold: float = 1e-3,
        rot_error_threshold: float = 0.1,
        run_lma_on_cpu: bool = True,
    ) -> Tuple[torch.Tensor, torch.Tensor]:
        assert target_poses.shape[1] == 7
        assert isinstance(repeat_counts, tuple)
        n_opt_steps_max = 3
        with torch.inference_mode():
            solutions, valids = self._generate_exact_ik_solutions(
                target_poses, repeat_counts[0], n_opt_steps_max, pos_error_threshold, rot_error_threshold, run_lma_on_cpu
            )
            if valids.all():
                return solutions, valids
            for i in range(1, len(repeat_counts)):
                missing = target_poses[torch.logical_not(valids), :]
                new_solutions, new_valids = self._generate_exact_ik_solutions(
                    missing, repeat_counts[i], n_opt_steps_max, pos_error_threshold, rot_error_threshold, run_lma_on_cpu
                )
                solutions[torch.logical_not(valids), :] = new_solutions
                valids[torch.logical_not(valids)] = new_valids
                if new_solutions.all():  # (sic) the reference tests the solution tensor, ikflow_solver.py:402
                    return solutions, valids
            return solutions, valids
