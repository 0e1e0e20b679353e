/* ikflow_b200.h -- C ABI of libikflow_b200.so, the B200 (sm_100a) engine behind the IKFlow hot path.
 *
 * The reference (jstmn/ikflow @ 2f4636e) is pure Python; its "FFI" for this path is the two operator
 * boundaries its solver calls into third-party Python packages:
 *
 *   (1) flow        ikflow/ikflow_solver.py:98    output_rev, _ = self.nn_model(latent, c=conditional, rev=True)
 *                   + :99-102 (slice [:, :ndof], robot.clamp_to_joint_limits)
 *   (2) kinematics  ikflow/ikflow_solver.py:114   robot.forward_kinematics(qs)
 *                   ikflow/ikflow_solver.py:116   geodesic_distance_between_quaternions(...)
 *                   ikflow/ikflow_solver.py:205,208  robot.inverse_kinematics_step_levenburg_marquardt(poses, q)
 *                   ikflow/ikflow_solver.py:201-233  the LM / select / compact loop of _generate_exact_ik_solutions
 *
 * Every entry point below replaces one of those calls; the reference-side binding (a ctypes stub) is shown in
 * INTEGRATION.md.  Conventions:
 *   - plain pointers and sizes only; all tensor pointers are DEVICE pointers owned by the caller (torch), row-major
 *     fp32 unless stated, borrowed for the duration of the call;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*; NULL = legacy default stream) and does
 *     no hidden host synchronisation (exceptions are documented);
 *   - return 0 on success, a negative IKF_E* code otherwise; never throws.  ikf_last_error() returns a thread-local
 *     human-readable message for the last failing call;
 *   - handles are re-entrant: any number of host threads and streams may share one.  A flow handle owns one exchange
 *     workspace and a launch occupies every SM, so the library serialises its launches (a mutex on the host side, an
 *     event wait when consecutive launches come from different streams); results are those of the calls in submission
 *     order.  Inside a CUDA stream capture use one capturing stream per handle.
 */
#ifndef IKFLOW_B200_H_
#define IKFLOW_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IKF_OK 0
#define IKF_EINVAL (-1)   /* bad argument */
#define IKF_ECUDA (-2)    /* CUDA runtime error (see ikf_last_error) */
#define IKF_ENOMEM (-3)   /* device allocation failed */
#define IKF_EDEVICE (-4)  /* device is not an sm_100 part / kernel image not loadable */
#define IKF_ESTATUS (-5)  /* an EARLIER launch on the handle reported IKF_STATUS_SYNC_TIMEOUT or IKF_STATUS_RANGE (its output is invalid) */

/* status bits reported by ikf_flow_status / ikf_flow_poll_status */
#define IKF_STATUS_NONFINITE 1u    /* an output was not finite (NaN / inf inputs propagate, as in the reference) */
#define IKF_STATUS_SYNC_TIMEOUT 2u /* an inter-CTA dependency wait timed out after ~1.3 s (results invalid): the teams of a
                                      launch spin on each other's flags, so a co-tenant kernel that holds SMs can stall
                                      it; the kernel then gives up instead of hanging the GPU, writes this bit into
                                      mapped host memory, and the NEXT call on the handle fails with IKF_ESTATUS */

#define IKF_STATUS_RANGE 4u        /* IKF_PRECISION_FP16X3 only: a hidden activation left the fp16 range (|x| > 65504): the
                                      outputs of that launch are invalid (NaN); like a timeout, the NEXT call on the handle
                                      fails with IKF_ESTATUS.  Use IKF_PRECISION_BF16X3 for such weights */

#define IKF_MAX_WIDTH 16  /* dim_latent_space */
#define IKF_MAX_LINKS 16  /* links on the kinematic chain (fixed + actuated) */
#define IKF_MAX_DOF 8     /* actuated joints handled by the kinematics kernels (one lane each) */

typedef struct IkfFlow IkfFlow;
typedef struct IkfRobot IkfRobot;

/* Hyper-parameters of the conditional flow -- the fields of ikflow/model_descriptions.yaml:10-17 plus what
 * IKFlowSolver.__init__ derives from them (ikflow/ikflow_solver.py:51-54). */
typedef struct {
  int32_t ndim_tot;        /* dim_latent_space (network width W), 2..IKF_MAX_WIDTH */
  int32_t dim_cond;        /* 8 with softflow (default), 7 without -- ikflow_solver.py:51-53 */
  int32_t nb_nodes;        /* number of [PermuteRandom, GLOWCouplingBlock] pairs -- model.py:338 */
  int32_t coeff_fn_config; /* n_layers of subnet_constructor (1..4): n_layers+1 nn.Linear -- model.py:51-96 */
  int32_t hidden;          /* coeff_fn_internal_size; multiple of 64 */
  int32_t ndof;            /* robot.ndof <= ndim_tot */
  float rnvp_clamp;        /* GLOWCouplingBlock clamp -- model.py:347 */
  int32_t precision;       /* IKF_PRECISION_*: operand format of the hidden-layer tensor-core products */
} IkfFlowDesc;

/* Hidden 1024x1024 layers run on the tensor cores with fp32 accumulation.  BF16X3 splits every fp32 operand into a
 * bf16 head and a bf16 tail and issues head*head + head*tail + tail*head (16 mantissa bits, fp32 exponent range:
 * no overflow/underflow hazards) -- 3e-5 abs from the fp32 reference after 96 chained layers (scripts/precision_study.py).
 * BF16X1 is the single-product fast mode (about 1e-2 abs): NOT parity grade, reported separately. */
#define IKF_PRECISION_BF16X3 0
#define IKF_PRECISION_BF16X1 1
/* FP16X3: fp16 head + fp16 tail scaled by 2^11 (22 mantissa bits), the two correction products in their own fp32
 * accumulator: as close to the fp32 reference as fp32 is to fp64 (5e-6 abs on the synthetic Panda weights, 1e-5 with the
 * last layers x2 where BF16X3 is at 2e-4) for the same three products.  Range of fp16: hidden activations must stay
 * below 65504 in magnitude (they are O(1) for a trained network); beyond it outputs are NaN and IKF_STATUS_NONFINITE
 * is raised.  tcgen05 engine only. */
#define IKF_PRECISION_FP16X3 2
/* AUTO: the most faithful format the engine of this model offers -- FP16X3 on the tcgen05 engine, BF16X3 on the mma.sync
 * engine (hidden sizes that are not a multiple of 128, or > 1024).  ikf_flow_precision() tells which. */
#define IKF_PRECISION_AUTO 3

/* ---- flow --------------------------------------------------------------------------------------------------------
 * ikf_flow_create: replaces glow_cNF_model(...) + nn_model.load_state_dict(...) (ikflow/model.py:291-356,
 * ikflow/ikflow_solver.py:413-429).  All inputs are HOST pointers, copied/repacked during the call (synchronous).
 *   weights   the nn.Linear parameters in state-dict order, fp32, concatenated:
 *             for i in 0..nb_nodes-1: for subnet in (subnet1, subnet2): for each Linear: weight [out,in] then bias [out]
 *             (keys module_list.{2+2i}.subnet{1,2}.{0,2,..}.{weight,bias})
 *   perm_inv  [nb_nodes][ndim_tot] int64, module_list.{1+2i}.perm_inv
 *   m_inv     [ndim_tot][ndim_tot] module_list.0.M_inv;  flt_b [ndim_tot] module_list.0.b
 *   joint_lo/joint_hi  [ndof] robot.actuated_joints_limits (for the clamp of ikflow_solver.py:102)
 */
int ikf_flow_create(const IkfFlowDesc* desc, const float* weights, size_t n_weights, const int64_t* perm_inv,
                    const float* m_inv, const float* flt_b, const float* joint_lo, const float* joint_hi, int device,
                    IkfFlow** out);
void ikf_flow_destroy(IkfFlow* flow);

/* Number of fp32 values ikf_flow_create expects in `weights` for `desc` (0 on invalid desc). */
size_t ikf_flow_weight_count(const IkfFlowDesc* desc);

/* ikf_flow_inverse: replaces ikflow_solver.py:98-102 -- the whole reverse pass glow_{nb-1}^-1, perm_{nb-1}^-1, ...,
 * glow_0^-1, perm_0^-1, FixedLinearTransform^-1, followed (optionally) by the joint-limit clamp.
 *   latent   [batch][ndim_tot]            (row stride latent_ld floats)
 *   cond     [cond_rows][cond_cols]       (row stride cond_ld floats); row b of the batch uses cond row (b % cond_rows):
 *            cond_rows == batch  -> one pose per row (ikflow_solver.py:338)
 *            cond_rows == 1      -> one pose broadcast (y.expand, ikflow_solver.py:334-336)
 *            cond_rows == n      -> conditional.repeat((repeat_count, 1)) of ikflow_solver.py:185
 *            cond_cols may be dim_cond or 7; missing trailing columns are 0 (the softflow column, ikflow_solver.py:335)
 *   out      [batch][out_cols]            (row stride out_ld floats); out_cols <= ndim_tot.  out_cols = ndim_tot gives
 *            nn_model's raw output, out_cols = ndof the solver's `solutions`
 *   clamp    != 0: clamp columns < ndof to the joint limits (ikflow_solver.py:101-102)
 */
int ikf_flow_inverse(IkfFlow* flow, const float* latent, int latent_ld, const float* cond, int cond_ld, int cond_rows,
                     int cond_cols, float* out, int out_ld, int out_cols, int batch, int clamp, void* stream);

/* The same pass restricted to coupling blocks block_first, block_first-1, ..., block_last (the order of the reverse
 * pass; nb_nodes-1 >= block_first >= block_last >= 0): state_out = perm_{last}^-1 glow_{last}^-1 ... glow_{first}^-1
 * (state_in).  One launch of the per-block kernel chain; used to compare against the reference block by block.
 * state_in/state_out are [batch][ndim_tot] (row strides in floats); FixedLinearTransform^-1 and the clamp are applied
 * only by ikf_flow_inverse. */
int ikf_flow_inverse_blocks(IkfFlow* flow, const float* state_in, int in_ld, const float* cond, int cond_ld,
                            int cond_rows, int cond_cols, float* state_out, int out_ld, int batch, int block_first,
                            int block_last, void* stream);

/* Forward pass (x -> z) with its log-determinant -- nn_model(x, c=cond, rev=False) of the reference
 * (ikflow/training/lt_model.py:129-159 ml_loss_fn; tests/model_test.py:117-177): FixedLinearTransform (x.mm(M) + b), then
 * for i = 0 .. nb_nodes-1: PermuteRandom_i (x[:, perm_i]) and GLOWCouplingBlock_i forward (subnet2 first).  Inference
 * only (no gradients).  tcgen05 engine only (IKF_EINVAL otherwise).
 *   x        [batch][ndim_tot]  (row stride x_ld floats)        cond as for ikf_flow_inverse
 *   z_out    [batch][ndim_tot]  (row stride out_ld floats)
 *   logdet_out [batch] or NULL: log|det dz/dx| = logDetM + sum over blocks of the clamped scales */
int ikf_flow_forward(IkfFlow* flow, const float* x, int x_ld, const float* cond, int cond_ld, int cond_rows, int cond_cols,
                     float* z_out, int out_ld, float* logdet_out, int batch, void* stream);

/* ---- fused gather of the batch-sharded solve (multi-GPU, one process per GPU) -------------------------------------------
 * The reference has no multi-GPU path; BASELINE's north star shards the batch over up to 8 GPUs and gathers the joint
 * angles with one NCCL all-gather.  The fused alternative removes the collective from the critical path: the kernel's
 * final epilogue stores its rows into the gathered buffer of EVERY rank (peer-mapped device pointers, NVLink stores) and
 * the last CTA to finish raises this rank's flag on every rank and then waits for the flags of the other ranks, so that
 * the end of the kernel IS the end of the gather: no second launch.
 *
 * ikf_flow_set_peers: gather_bufs[r] = base of rank r's symmetric allocation as mapped into THIS process (cudaIpc /
 * torch symmetric memory / cuMem), flag_bufs[r] = rank r's flag array [n_ranks] uint32 (zero-initialised by its owner
 * before the first call; all ranks must have called set_peers before any rank launches).  n_ranks = 0 switches it off.
 * Every rank must then make the same sequence of ikf_flow_inverse_gather calls (SPMD). */
int ikf_flow_set_peers(IkfFlow* flow, int n_ranks, int rank, void* const* gather_bufs, void* const* flag_bufs);

/* ikf_flow_inverse with the gather fused in: this rank's `batch` rows (>= 1) become rows [row0, row0 + batch) of the
 * gathered tensor [rows_total][gather_ld] that starts `gather_offset` floats into every rank's symmetric allocation.
 * When the call's work has completed on `stream`, the gathered tensor of THIS rank holds the rows of all ranks.
 * The caller alternates between two offsets (ping-pong) so that a rank that runs ahead never overwrites a buffer a
 * peer is still reading. */
int ikf_flow_inverse_gather(IkfFlow* flow, const float* latent, int latent_ld, const float* cond, int cond_ld, int cond_rows,
                            int cond_cols, int out_cols, int batch, int clamp, size_t gather_offset, int gather_ld, int row0,
                            void* stream);

/* Optional: the stored FixedLinearTransform parameters M [ndim_tot][ndim_tot] (row-major, "module_list.0.M", HOST
 * pointer) and logDetM for the forward pass.  Without this call ikf_flow_create's own M = M_inv^-1 (double precision
 * Gauss-Jordan) and log|det M| are used. */
int ikf_flow_set_forward_tables(IkfFlow* flow, const float* m, float log_det_m);

/* Reads (and clears) the device status word.  SYNCHRONISES `stream`. */
int ikf_flow_status(IkfFlow* flow, void* stream, uint32_t* status_out);

/* The same bits WITHOUT synchronising: reads the mirror the kernels keep in mapped host memory (complete for every
 * launch whose stream has been synchronised; may already show a launch still in flight).  Does not clear. */
int ikf_flow_poll_status(IkfFlow* flow, uint32_t* status_out);

/* The operand format in use (IKF_PRECISION_*; resolves IKF_PRECISION_AUTO). */
int ikf_flow_precision(IkfFlow* flow);

/* Name of the kernel the last launch on this handle used (for benchmark / profile bookkeeping). */
const char* ikf_flow_last_kernel(IkfFlow* flow);
/* CTAs per thread-block cluster of the last launch (1 = no clusters): the CTAs holding the same weight slice of
 * neighbouring teams load every weight chunk once and multicast it (IKFLOW_B200_CLUSTER=1|2|4 overrides the default). */
int ikf_flow_last_cluster(IkfFlow* flow);

/* Introspection for benchmarks: bytes of packed weights resident in HBM, CTAs per launch, dynamic smem per CTA. */
int ikf_flow_info(IkfFlow* flow, size_t* packed_weight_bytes, int* grid_ctas_last, int* smem_bytes);

/* Debug aid: the CTAs of team 0 write %globaltimer stamps [cta][layer][96] (layer = 4*subnet step + layer index) into
 * dev_stamps (device memory, (hidden/64)*n_layers*96 uint64) during the next launches.  n_layers = 0 switches it off. */
int ikf_flow_debug_trace(IkfFlow* flow, unsigned long long* dev_stamps, int n_layers);

/* ---- kinematics --------------------------------------------------------------------------------------------------
 * ikf_robot_create: replaces jrl.robots.get_robot(name) for the functions on the hot path.  HOST pointers.
 *   kind      [n_links] 0 = fixed, 1 = revolute, 2 = prismatic (chain order base -> end effector)
 *   fixed_T   [n_links][12] row-major 3x4 [R | p] of the joint's origin (URDF xyz/rpy), fp64
 *   axis      [n_links][3] joint axis in the joint frame (ignored for fixed)
 *   lo, hi    [ndof] limits of the actuated joints in chain order
 */
int ikf_robot_create(int n_links, const int32_t* kind, const double* fixed_T, const double* axis, const double* lo,
                     const double* hi, int device, IkfRobot** out);
void ikf_robot_destroy(IkfRobot* robot);
int ikf_robot_ndof(const IkfRobot* robot);

/* robot.forward_kinematics(q[m,ndof]) -> poses[m,7] = [x y z qw qx qy qz]  (ikflow_solver.py:114) */
int ikf_forward_kinematics(IkfRobot* robot, const float* q, float* poses_out, int m, void* stream);

/* robot.sample_joint_angles_and_poses(n) minus the self-collision filter (jrl; called at scripts/benchmark_runtime.py:83-86,
 * scripts/evaluate.py:137-139, tests/ikflow_solver_test.py:70-72): q[s][j] ~ U(lo_j + eps, hi_j - eps), poses = FK(q), one
 * launch.  Counter-based Philox4x32-10: key = seed, counter = (first_index + s, block j / 4), word j % 4,
 * u = (word + 0.5) * 2^-32, q = fp32(lo' + u * (hi' - lo')) evaluated in fp64.  Sample s depends only on
 * (seed, first_index + s), so shards of one index range are independent of how it is split over calls or GPUs.
 *   q_out [m][ndof], poses_out [m][7] */
int ikf_sample_joint_angles_and_poses(IkfRobot* robot, uint64_t seed, uint64_t first_index, double joint_limit_eps,
                                      float* q_out, float* poses_out, int m, void* stream);

/* robot.clamp_to_joint_limits(q) in place (ikflow_solver.py:102); NaN stays NaN (torch.clamp) */
int ikf_clamp_to_joint_limits(IkfRobot* robot, float* q, int m, void* stream);

/* robot.inverse_kinematics_step_levenburg_marquardt(target_poses, q) (ikflow_solver.py:205,208):
 * J = jacobian(q); e = [rpy(q_t (x) q_cur^-1); p_t - p_cur]; dq = solve(J^T J + lambd I, J^T e); q' = clamp(q + dq).
 * Sample i uses pose row (i % pose_rows).  q_out may alias q_in. */
int ikf_lm_step(IkfRobot* robot, const float* poses, int pose_rows, const float* q_in, float* q_out, int m,
                float lambd, int clamp, void* stream);

/* IKFlowSolver._calculate_pose_error (ikflow_solver.py:112-117): L2 position error and quaternion geodesic. */
int ikf_pose_error(IkfRobot* robot, const float* q, const float* poses, int pose_rows, float* pos_err, float* rot_err,
                   int m, void* stream);

/* The LM / select / compact loop of _generate_exact_ik_solutions (ikflow_solver.py:197-233), device-side:
 *   q_seeds   [repeat_count * n][ndof]  flow samples, row k*n + p = repeat k of pose p (ikflow_solver.py:185); updated
 *             in place (each row holds its iterate at the first step it became valid, or after n_steps)
 *   poses     [n][7]
 *   final_q   [n][ndof], final_valid [n] (uint8): for every pose the LAST valid repeat at the FIRST step where any
 *             repeat is valid (the overwrite order of the Python loop at :217-222); rows of unsolved poses are 0
 *   n_valid_dev  optional device int32: number of solved poses
 */
int ikf_lm_refine(IkfRobot* robot, const float* poses, float* q_seeds, int n, int repeat_count, int n_steps,
                  float pos_thr, float rot_thr, float lambd, float* final_q, uint8_t* final_valid,
                  int32_t* n_valid_dev, void* stream);

/* evaluate_solutions minus self-collision (ikflow/evaluation_utils.py:65-112,130-147): pose errors + joint-limit flag */
int ikf_evaluate_solutions(IkfRobot* robot, const float* q, const float* poses, int pose_rows, float* pos_err,
                           float* rot_err, uint8_t* limits_exceeded, int m, void* stream);

/* ---- misc -------------------------------------------------------------------------------------------------------- */
const char* ikf_last_error(void);
const char* ikf_version(void);
/* Number of kernels this library has launched in this process (for bench.py's gpu_launches claim). */
uint64_t ikf_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* IKFLOW_B200_H_ */
