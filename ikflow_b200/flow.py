"""``FlowModel`` -- the object the solver stores in ``nn_model``; drop-in for the FrEIA ``GraphINN`` at the one call the
hot path makes:  ``output_rev, _ = self.nn_model(latent, c=conditional, rev=True)``  (``ikflow/ikflow_solver.py:98``).

It owns a handle of the sm_100a flow engine (``csrc/flow.cu``) per CUDA device, loads state dicts with the FrEIA key
names, and has no CPU execution path.
"""

import ctypes
from typing import Dict, Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .model import IkflowModelParameters, state_dict_keys, subnet_layer_dims


class FlowModel:
    def __init__(
        self,
        params: IkflowModelParameters,
        joint_limits: Sequence[Tuple[float, float]],
        dim_cond: int,
        ndim_tot: int,
        precision: str = "auto",
    ):
        assert params.coupling_layer == "glow", "only the GLOW coupling block is implemented (all released models)"
        assert not getattr(params, "sigmoid_on_output", False), "sigmoid_on_output is not used by any released model"
        assert params.permute_random_enabled, "permute_random_enabled=False is not supported"
        assert precision in ("auto", "bf16x3", "bf16x1", "fp16x3")
        self.params = params
        self.ndim_tot = int(ndim_tot)
        self.dim_cond = int(dim_cond)
        self.ndof = len(joint_limits)
        self.joint_limits = [(float(lo), float(hi)) for lo, hi in joint_limits]
        self.nb_nodes = int(params.nb_nodes)
        self.coeff_fn_config = int(params.coeff_fn_config)
        self.hidden = int(params.coeff_fn_internal_size)
        self.rnvp_clamp = float(params.rnvp_clamp)
        self.precision = precision
        self._state_dict: Optional[Dict[str, torch.Tensor]] = None
        self._handles: Dict[int, int] = {}
        self._expected = state_dict_keys(params_with_width(params, self.ndim_tot), self.dim_cond)

    # ---- torch.nn.Module-like surface the reference touches ---------------------------------------------------------
    def eval(self):
        return self

    def train(self, mode: bool = True):
        assert not mode, "ikflow_b200 is inference only"
        return self

    def to(self, *args, **kwargs):
        return self

    def parameters(self) -> Iterator[torch.Tensor]:
        return iter(self._state_dict.values()) if self._state_dict else iter(())

    def state_dict(self) -> Dict[str, torch.Tensor]:
        assert self._state_dict is not None, "no state dict loaded"
        return dict(self._state_dict)

    def load_state_dict(self, state_dict: Dict[str, torch.Tensor], strict: bool = True):
        """Accepts the FrEIA key layout (``module_list.{k}...``), optionally prefixed by ``_orig_mod.`` (what the
        reference adds for compiled models, ``ikflow_solver.py:420-426``) or ``nn_model.`` (Lightning checkpoints,
        ``scripts/download_model_from_wandb_checkpoint.py:13-28``)."""
        sd = {}
        for k, v in state_dict.items():
            for prefix in ("_orig_mod.", "nn_model."):
                if k.startswith(prefix):
                    k = k[len(prefix):]
            sd[k] = v.detach().cpu() if isinstance(v, torch.Tensor) else torch.as_tensor(v)
        missing = [k for k in self._expected if k not in sd]
        unexpected = [k for k in sd if k not in self._expected]
        if missing or (strict and unexpected):
            raise RuntimeError(
                f"Error(s) in loading state_dict for FlowModel: missing keys {missing[:4]}{'...' if len(missing) > 4 else ''}"
                f", unexpected keys {unexpected[:4]}{'...' if len(unexpected) > 4 else ''}"
            )
        for k, shape in self._expected.items():
            if tuple(sd[k].shape) != tuple(shape):
                raise RuntimeError(f"size mismatch for {k}: checkpoint {tuple(sd[k].shape)}, model {tuple(shape)}")
        self._state_dict = {k: sd[k] for k in self._expected}
        for h in self._handles.values():
            _lib.lib().ikf_flow_destroy(h)
        self._handles = {}

    def __del__(self):
        try:
            for h in self._handles.values():
                _lib.lib().ikf_flow_destroy(h)
        except Exception:
            pass

    # ---- engine -----------------------------------------------------------------------------------------------------
    def _desc(self) -> _lib.IkfFlowDesc:
        return _lib.IkfFlowDesc(
            self.ndim_tot, self.dim_cond, self.nb_nodes, self.coeff_fn_config, self.hidden, self.ndof, self.rnvp_clamp,
            {"bf16x3": _lib.IKF_PRECISION_BF16X3, "bf16x1": _lib.IKF_PRECISION_BF16X1, "fp16x3": _lib.IKF_PRECISION_FP16X3, "auto": _lib.IKF_PRECISION_AUTO}[self.precision],
        )

    def flat_weights(self) -> np.ndarray:
        """The nn.Linear parameters in the order ``ikf_flow_create`` expects (state-dict order)."""
        sd = self._state_dict
        parts: List[np.ndarray] = []
        n_linear = self.coeff_fn_config + 1
        for i in range(self.nb_nodes):
            for sub in ("subnet1", "subnet2"):
                for li in range(n_linear):
                    p = f"module_list.{2 + 2 * i}.{sub}.{2 * li}"
                    parts.append(sd[p + ".weight"].to(torch.float32).contiguous().numpy().reshape(-1))
                    parts.append(sd[p + ".bias"].to(torch.float32).contiguous().numpy().reshape(-1))
        return np.ascontiguousarray(np.concatenate(parts))

    def _handle(self, device: torch.device) -> int:
        if self._state_dict is None:
            raise RuntimeError("FlowModel has no weights: call load_state_dict(...) first")
        idx = device.index if device.index is not None else torch.cuda.current_device()
        if idx not in self._handles:
            sd = self._state_dict
            weights = self.flat_weights()
            perm_inv = np.ascontiguousarray(
                np.stack([sd[f"module_list.{1 + 2 * i}.perm_inv"].numpy().astype(np.int64) for i in range(self.nb_nodes)])
            )
            m_inv = np.ascontiguousarray(sd["module_list.0.M_inv"].to(torch.float32).numpy())
            flt_b = np.ascontiguousarray(sd["module_list.0.b"].to(torch.float32).numpy().reshape(-1))
            lo = np.array([l for l, _ in self.joint_limits], dtype=np.float32)
            hi = np.array([h for _, h in self.joint_limits], dtype=np.float32)
            desc = self._desc()
            out = ctypes.c_void_p()
            code = _lib.lib().ikf_flow_create(
                ctypes.byref(desc), weights.ctypes.data, weights.size, perm_inv.ctypes.data, m_inv.ctypes.data,
                flt_b.ctypes.data, lo.ctypes.data, hi.ctypes.data, idx, ctypes.byref(out),
            )
            _lib.check(code, "ikf_flow_create")
            # forward pass: the stored FixedLinearTransform parameters instead of the library's own M_inv^-1
            m = np.ascontiguousarray(sd["module_list.0.M"].to(torch.float32).numpy())
            _lib.check(_lib.lib().ikf_flow_set_forward_tables(out, m.ctypes.data, float(sd["module_list.0.logDetM"])), "ikf_flow_set_forward_tables")
            self._handles[idx] = out.value
        return self._handles[idx]

    @staticmethod
    def _check_inputs(latent: torch.Tensor, cond: torch.Tensor, width: int, max_cond: int):
        for name, t in (("latent", latent), ("conditional", cond)):
            if not isinstance(t, torch.Tensor):
                raise TypeError(f"{name} must be a torch.Tensor (got {type(t)})")
            if not t.is_cuda:
                raise RuntimeError(
                    f"{name} is on '{t.device}': ikflow_b200 computes on CUDA (sm_100a) only, there is no CPU path"
                )
            assert t.dtype == torch.float32, f"{name} must be float32 (got {t.dtype})"
            assert t.dim() == 2, f"{name} must be 2-dimensional"
        assert latent.shape[1] == width, f"latent must be [n x {width}] (got {tuple(latent.shape)})"
        assert 7 <= cond.shape[1] <= max_cond, f"conditional must have 7..{max_cond} columns (got {cond.shape[1]})"
        assert latent.device == cond.device

    def inverse(
        self, latent: torch.Tensor, cond: torch.Tensor, out_cols: Optional[int] = None, clamp: bool = False
    ) -> torch.Tensor:
        """The whole reverse pass in one launch.  ``cond`` may have 1 row (broadcast), ``latent.shape[0]`` rows, or
        n rows with ``latent.shape[0] = r*n`` (repeat-major tiling, ``conditional.repeat((r, 1))``)."""
        self._check_inputs(latent, cond, self.ndim_tot, self.dim_cond)
        batch = latent.shape[0]
        assert cond.shape[0] >= 1 and batch % cond.shape[0] == 0, f"{batch} rows vs {cond.shape[0]} condition rows"
        out_cols = self.ndim_tot if out_cols is None else out_cols
        latent, cond = latent.contiguous(), cond.contiguous()
        out = torch.empty((batch, out_cols), dtype=torch.float32, device=latent.device)
        code = _lib.lib().ikf_flow_inverse(
            self._handle(latent.device), latent.data_ptr(), latent.stride(0), cond.data_ptr(), cond.stride(0),
            cond.shape[0], cond.shape[1], out.data_ptr(), out.stride(0), out_cols, batch, int(clamp),
            torch.cuda.current_stream(latent.device).cuda_stream,
        )
        _lib.check(code, "ikf_flow_inverse")
        return out

    # ---- fused gather of the batch-sharded solve (see ikflow_b200/distributed.py: PeerGather) -------------------------------
    def set_peers(self, device: torch.device, n_ranks: int, rank: int, gather_ptrs: Sequence[int], flag_ptrs: Sequence[int]) -> None:
        """``ikf_flow_set_peers``: peer-mapped base pointers of every rank's gathered buffer and flag array."""
        bufs = (ctypes.c_void_p * max(n_ranks, 1))(*[ctypes.c_void_p(int(x)) for x in gather_ptrs])
        flags = (ctypes.c_void_p * max(n_ranks, 1))(*[ctypes.c_void_p(int(x)) for x in flag_ptrs])
        _lib.check(_lib.lib().ikf_flow_set_peers(self._handle(device), n_ranks, rank, bufs, flags), "ikf_flow_set_peers")

    def inverse_gather(self, latent: torch.Tensor, cond: torch.Tensor, out_cols: int, clamp: bool, gather_offset: int, gather_ld: int, row0: int) -> None:
        """The reverse pass of this rank's rows with the gather fused into the kernel's epilogue (``ikf_flow_inverse_gather``):
        rows [row0, row0 + n) of the gathered tensor at ``gather_offset`` floats inside every rank's symmetric buffer."""
        self._check_inputs(latent, cond, self.ndim_tot, self.dim_cond)
        batch = latent.shape[0]
        assert cond.shape[0] >= 1 and batch % cond.shape[0] == 0, f"{batch} rows vs {cond.shape[0]} condition rows"
        latent, cond = latent.contiguous(), cond.contiguous()
        code = _lib.lib().ikf_flow_inverse_gather(
            self._handle(latent.device), latent.data_ptr(), latent.stride(0), cond.data_ptr(), cond.stride(0), cond.shape[0],
            cond.shape[1], out_cols, batch, int(clamp), gather_offset, gather_ld, row0, torch.cuda.current_stream(latent.device).cuda_stream,
        )
        _lib.check(code, "ikf_flow_inverse_gather")

    def forward_pass(self, x: torch.Tensor, cond: torch.Tensor):
        """x -> z with its log-determinant, ``nn_model(x, c=cond, rev=False)`` (``ikflow/training/lt_model.py:156``), in one
        launch.  Inference only: no gradients flow through it."""
        self._check_inputs(x, cond, self.ndim_tot, self.dim_cond)
        batch = x.shape[0]
        assert cond.shape[0] >= 1 and batch % cond.shape[0] == 0, f"{batch} rows vs {cond.shape[0]} condition rows"
        x, cond = x.contiguous(), cond.contiguous()
        z = torch.empty((batch, self.ndim_tot), dtype=torch.float32, device=x.device)
        logdet = torch.empty((batch,), dtype=torch.float32, device=x.device)
        code = _lib.lib().ikf_flow_forward(
            self._handle(x.device), x.data_ptr(), x.stride(0), cond.data_ptr(), cond.stride(0), cond.shape[0], cond.shape[1],
            z.data_ptr(), z.stride(0), logdet.data_ptr(), batch, torch.cuda.current_stream(x.device).cuda_stream,
        )
        _lib.check(code, "ikf_flow_forward")
        return z, logdet

    def inverse_blocks(self, state: torch.Tensor, cond: torch.Tensor, block_first: int, block_last: int) -> torch.Tensor:
        """Coupling blocks block_first, block_first-1, ..., block_last of the reverse pass (each followed by its
        permutation); no FixedLinearTransform, no clamp."""
        self._check_inputs(state, cond, self.ndim_tot, self.dim_cond)
        state, cond = state.contiguous(), cond.contiguous()
        out = torch.empty_like(state)
        code = _lib.lib().ikf_flow_inverse_blocks(
            self._handle(state.device), state.data_ptr(), state.stride(0), cond.data_ptr(), cond.stride(0),
            cond.shape[0], cond.shape[1], out.data_ptr(), out.stride(0), state.shape[0], block_first, block_last,
            torch.cuda.current_stream(state.device).cuda_stream,
        )
        _lib.check(code, "ikf_flow_inverse_blocks")
        return out

    def status(self, device=None) -> int:
        """Reads and clears the engine's status word (synchronises the current stream)."""
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        word = ctypes.c_uint32(0)
        code = _lib.lib().ikf_flow_status(self._handle(dev), torch.cuda.current_stream(dev).cuda_stream, ctypes.byref(word))
        _lib.check(code, "ikf_flow_status")
        return int(word.value)

    def poll_status(self, device=None) -> int:
        """The status bits WITHOUT synchronising (``ikf_flow_poll_status``: a mirror the kernels keep in mapped host
        memory).  A launch that timed out (``IKF_STATUS_SYNC_TIMEOUT``) additionally makes the NEXT call on the handle
        raise ``IkflowB200Error`` (code ``IKF_ESTATUS``)."""
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        word = ctypes.c_uint32(0)
        _lib.check(_lib.lib().ikf_flow_poll_status(self._handle(dev), ctypes.byref(word)), "ikf_flow_poll_status")
        return int(word.value)

    def effective_precision(self, device=None) -> str:
        """The operand format in use: resolves ``"auto"`` (fp16x3 on the tcgen05 engine, bf16x3 on the mma.sync engine)."""
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        return {0: "bf16x3", 1: "bf16x1", 2: "fp16x3"}[int(_lib.lib().ikf_flow_precision(self._handle(dev)))]

    def last_kernel(self, device=None) -> str:
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        return _lib.lib().ikf_flow_last_kernel(self._handle(dev)).decode()

    def last_cluster(self, device=None) -> int:
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        return int(_lib.lib().ikf_flow_last_cluster(self._handle(dev)))

    def info(self, device=None) -> Dict[str, int]:
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        nbytes, grid, smem = ctypes.c_size_t(0), ctypes.c_int(0), ctypes.c_int(0)
        _lib.check(_lib.lib().ikf_flow_info(self._handle(dev), ctypes.byref(nbytes), ctypes.byref(grid), ctypes.byref(smem)), "ikf_flow_info")
        return {"packed_weight_bytes": nbytes.value, "grid_ctas_last": grid.value, "smem_bytes": smem.value}

    def __call__(self, x_or_z: torch.Tensor, c=None, rev: bool = False, jac: bool = True):
        """``GraphINN.forward``.  Returns ``(out [n x ndim_tot], logdet)``.  ``rev=True`` (what the solver uses,
        ``ikflow_solver.py:98``): the log-determinant is not computed (the solver discards it) and is returned as
        ``None``.  ``rev=False``: z and log|det dz/dx| (``None`` with ``jac=False``), inference only."""
        if isinstance(c, (list, tuple)):
            assert len(c) == 1
            c = c[0]
        assert c is not None, "the flow is conditional: pass c=[n x dim_cond]"
        assert x_or_z.shape[0] == c.shape[0], f"{x_or_z.shape[0]} != {c.shape[0]}"
        if not rev:
            z, logdet = self.forward_pass(x_or_z, c)
            return z, (logdet if jac else None)
        return self.inverse(x_or_z, c), None

    forward = __call__


def params_with_width(params: IkflowModelParameters, width: int) -> IkflowModelParameters:
    """Copy of ``params`` whose ``dim_latent_space`` is the network width actually built (the reference passes the
    width separately, ``ikflow/ikflow_solver.py:58``)."""
    p = IkflowModelParameters()
    p.__dict__.update(params.__dict__)
    p.dim_latent_space = width
    return p
