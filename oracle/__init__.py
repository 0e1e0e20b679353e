"""oracle/ -- CPU restatement of the jstmn/ikflow hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s baseline legs may import this
package.  The product (``ikflow_b200``) never imports, links or executes anything in here; it
fails loudly when its CUDA library is missing.

What is restated (plain PyTorch fp32 ops, same op sequence as the reference's L2 dependencies):

* ``freia_flow.py``      FrEIA 0.2 reverse pass used at ``ikflow/ikflow_solver.py:98``
                         (GraphINN -> GLOWCouplingBlock / PermuteRandom / FixedLinearTransform as
                         wired by ``ikflow/model.py:291-356``).
* ``jrl_kinematics.py``  jrl @ 2ba7c39 ``Robot.forward_kinematics / jacobian /
                         inverse_kinematics_step_levenburg_marquardt / clamp_to_joint_limits`` and
                         ``math_utils.geodesic_distance_between_quaternions`` as called from
                         ``ikflow/ikflow_solver.py:102,114,116,205,208``.
* ``solver.py``          the host logic of ``ikflow/ikflow_solver.py`` itself (``_run_inference``,
                         ``_generate_exact_ik_solutions`` incl. the "last valid repeat wins" Python
                         loop at ``:217-222``, the retry driver at ``:345-411``).

PINNING STATUS
--------------
FrEIA 0.2 (PyPI sdist, ``uv.lock:533-541``) and jrl (git pin ``pyproject.toml:22``) are third-party
dependencies that are NOT vendored in ``/root/reference`` and are not installable here (no network,
not in ``/opt/wheelhouse``).  ``import ikflow`` itself dies at ``ikflow/config.py:6``.  Therefore:

* kinematics: PINNED against every known-answer vector the reference's own tests hold
  (``tests/evaluation_utils_test.py:20-32,37-55``, ``tests/model_test.py:18-44``) --
  see ``tests/test_oracle_kats.py``.
* flow arithmetic (GLOW coupling / permutation / fixed linear transform): **parity unpinned** --
  the reference holds no numeric golden vector for it (only the relational checks at
  ``tests/ikflow_solver_test.py:94-117``, which the oracle and the CUDA path both pass).  The
  restatement follows the published FrEIA 0.2 algorithm and the constraints visible at the
  reference's call sites (state-dict key names ``scripts/download_model_from_wandb_checkpoint.py:14-18``,
  the split_len note ``ikflow/model.py:320-336``, the in-repo copy of FixedLinearTransform
  ``ikflow/model.py:153-238``).
"""
