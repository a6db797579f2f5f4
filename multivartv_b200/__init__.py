"""multivartv_b200 -- B200-native (sm_100a) implementation of the MultivarTV mesh-based ADMM hot path.

The numerical work lives in ``lib/libmvtv_b200.so`` (hand-written CUDA behind the C ABI of
``include/mvtv.h``); this package is the thin host-side mirror of the reference's solver interface.
Importing the package does not load the library; the first call does, and fails loudly if it is missing.
"""
from . import _lib  # noqa: F401
from ._lib import (F32, F64, MODE_CPP, MODE_PY, MODE_RCPP, PRECOND_AUTO, PRECOND_CHEB1, PRECOND_CHEB2, PRECOND_CHEB3, PRECOND_CHEB4,  # noqa: F401
                   PRECOND_JACOBI, VARIANT_INTENDED,
                   VARIANT_REFERENCE, WARM_THETA_FROM_PLAN, WARM_U_FROM_PLAN, MvtvError, NotConverged)
from .solvers import (Plan, adapt_step, axes_from_mesh, create_deltas, create_lambdas, create_mesh, kfoldinds, mbs, mbs_mse,  # noqa: F401
                      mbs_one, mbs_predict,
                      mesh_axes, mesh_from_axes, nccl_unique_id, nearest1, pinned_empty, softthresh)

__all__ = ["Plan", "mbs", "create_lambdas", "kfoldinds", "mbs_one", "mbs_predict", "mbs_mse", "softthresh", "nearest1", "create_mesh", "mesh_axes",
           "mesh_from_axes", "axes_from_mesh", "create_deltas", "pinned_empty", "adapt_step", "MvtvError", "NotConverged"]
