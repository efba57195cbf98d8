"""
`FalkonOptions` look-alike (the reference builds it at
/root/reference/sobolev_alignment/krr_approx.py:267 from the user's `falkon_options` dict, and
/root/reference/sobolev_alignment/krr_model_selection.py:55-64 passes `never_store_kernel`,
`min_cuda_iter_size_32/64`, `num_fmm_streams`, ...).

Every upstream option is accepted.  The ones that steer Falkon's CPU/GPU block scheduler have no
meaning for a GPU-resident fused kernel and are stored but ignored; the solver constants
(`pc_epsilon_32`, `cg_epsilon_32`, `cg_tolerance`, `cg_full_gradient_every`) are honoured.  Unknown
keys raise a warning, never an error.
"""

from __future__ import annotations

import warnings

_DEFAULTS = {
    # solver constants (honoured)
    "pc_epsilon_32": 1e-5,
    "pc_epsilon_64": 1e-13,
    "cg_epsilon_32": 1e-7,
    "cg_epsilon_64": 1e-15,
    "cg_tolerance": 1e-7,
    "cg_full_gradient_every": 10,
    "cg_differential_convergence": False,
    # accepted for compatibility, no effect on the B200 path
    "debug": False,
    "use_cpu": False,
    "max_gpu_mem": float("inf"),
    "max_cpu_mem": float("inf"),
    "compute_arch_speed": False,
    "no_single_kernel": True,
    "min_cuda_pc_size_32": 10000,
    "min_cuda_pc_size_64": 30000,
    "min_cuda_iter_size_32": 300000000,
    "min_cuda_iter_size_64": 900000000,
    "never_store_kernel": False,
    "store_kernel_d_threshold": 1200,
    "num_fmm_streams": 2,
    "memory_slack": 0.9,
    "keops_acc_dtype": "auto",
    "keops_sum_scheme": "auto",
    "keops_active": "auto",
    "keops_memory_slack": 0.7,
    "chol_force_in_core": False,
    "chol_force_ooc": False,
    "chol_par_blk_multiplier": 2,
    "pc_epsilon": None,
    "cpu_preconditioner": False,
    "lauum_par_blk_multiplier": 8,
    # extensions of this implementation
    "row_sharded": False,      # X, Y passed to fit() are this rank's row shard (torch.distributed)
    "center_seed": None,       # seed of the centre draw when Falkon(seed=None); required if row_sharded
    "stage_rows": 65536,       # rows per pinned host->device staging chunk
    "device": None,            # cuda device index; default: current device
    "kernel_blocks": "recompute",  # "store": form K once per CG product pair through a transient row-block scratch
                                   # (design B; ignored when never_store_kernel=True)
    "kstore_budget": 24 << 30,     # bytes of that scratch
}


class FalkonOptions:
    def __init__(self, **kwargs):
        for k, v in _DEFAULTS.items():
            setattr(self, k, v)
        for k, v in kwargs.items():
            if k not in _DEFAULTS:
                warnings.warn(f"FalkonOptions: unknown option {k!r} is stored but ignored", stacklevel=2)
            setattr(self, k, v)
        if self.use_cpu:
            warnings.warn("FalkonOptions(use_cpu=True) is ignored: this implementation has no CPU path",
                          stacklevel=2)

    def pc_eps(self) -> float:
        return float(self.pc_epsilon) if self.pc_epsilon is not None else float(self.pc_epsilon_32)

    def __repr__(self):
        changed = {k: getattr(self, k) for k in _DEFAULTS if getattr(self, k) != _DEFAULTS[k]}
        return f"FalkonOptions({changed})"
