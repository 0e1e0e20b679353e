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

PINNING STATUS  (tests: ``tests/test_oracle_kats.py``, ``tests/test_reference_host_logic.py``)
--------------
FrEIA 0.2 (PyPI sdist, ``uv.lock:533-541``) and jrl (git pin ``pyproject.toml:22``) are third-party
dependencies that are NOT vendored in ``/root/reference`` and are not installable here (no network,
not in ``/opt/wheelhouse``); ``import ikflow`` itself dies at ``ikflow/config.py:6``.  Everything else on
the path is the reference's own Python and IS executed here, unmodified, from ``/root/reference`` on top
of stand-in ``jrl`` / ``FrEIA`` modules (``oracle/ref_stub.py``); its outputs at BASELINE sizes are frozen in
``tests/golden/reference_*.npz`` (``scripts/make_golden_reference.py``) for the GPU box, where the
reference tree does not exist.

| oracle function                                   | pinned by                                                        |
|---------------------------------------------------|------------------------------------------------------------------|
| jrl_kinematics.forward_kinematics (Panda)         | reference KAT ``tests/evaluation_utils_test.py:20-24``            |
| jrl_kinematics.pose_error / geodesic distance     | reference KAT ``tests/evaluation_utils_test.py:26-32`` (one point)|
| jrl_kinematics joint limits / limits-exceeded     | ``tests/model_test.py:18-44``, ``tests/evaluation_utils_test.py:37-55``; bit-equal to the reference's ``calculate_joint_limits_exceeded`` / ``solution_pose_errors`` run here |
| freia_flow.subnet_forward                         | bit-equal to the module ``ikflow.model.subnet_constructor`` (:51-96) builds, n_layers 1..4 |
| FixedLinearTransform (build_fixed_linear_transform + the last step of flow_inverse / first of flow_forward) | bit-equal (parameters, forward, reverse, log-det) to ``ikflow.model.IkFlowFixedLinearTransform`` (:153-238), the reference's documented copy of FrEIA's class; that class passes the reference's scaling KATs (``tests/model_test.py:50-106``) here |
| permute_random_tables                             | ``np.random.seed(i); np.random.permutation(W)`` itself + SURVEY App. F |
| flow_inverse / flow_forward WIRING (block order, seeds, ``split_len = W // 2``, ``M``, state-dict keys and shapes, which subnet feeds which half) | bit-equal to the graph ``ikflow.model.glow_cNF_model`` (:291-356) wires, for panda / fetch_arm / odd-width / no-softflow geometries -- on the stand-in FrEIA classes, whose subnets are the reference's own ``nn.Sequential`` |
| solver.OracleSolver (``_run_inference``, ``generate_ik_solutions``, ``_generate_exact_ik_solutions``, ``generate_exact_ik_solutions``, ``draw_latent``) | bit-equal to ``ikflow.ikflow_solver.IKFlowSolver`` run here: all branches (early return, r = 3 / r = 10 retries, never-converging poses, ``run_lma_on_cpu`` True/False below and above n = 750), n up to 2048 |
| **GLOW coupling formula** (``glow_coupling_reverse / _forward``: ``s = clamp * 0.636 * atan(a)``, ``y = (x - t) * exp(-s)``, subnet/half pairing) | **UNPINNED** -- lives in FrEIA 0.2 only; the reference holds no numeric vector for it (relational checks ``tests/ikflow_solver_test.py:94-117`` pass).  Self-consistency only: log-det = autograd Jacobian, forward(reverse) = id.  The note at ``ikflow/model.py:320-336`` says the released weights were trained on a pre-2021 FrEIA; whether that version clamped as ``exp(clamp*0.636*atan(s/clamp))`` must be re-checked against a FrEIA 0.2 sdist or a released ``.pkl`` (FK error of approximate solutions would expose it) the moment either is available |
| jrl LM step (``lm_step``: error vector layout, ``J^T J + lambda I``, clamp) and Jacobian row order, FetchArm / Fetch chain constants | jrl only; functionally pinned by the reference's closure assertions (``tests/ikflow_solver_test.py:82-87``: after refinement pos <= 1 mm, rot < 0.01 rad, solution == its own clamp), which hold for every pose the oracle marks valid |
"""
