"""cvxopt_b200 — B200-native KKT hot path for CVXOPT's cone solvers.

Only what the path needs: the C-ABI CUDA library (csrc/, libcvxopt_b200.so) and
the host-side mirror of the reference's kktsolver / misc_solvers interface.
"""
from ._lib import load, exported_symbols, LIB_PATH  # noqa: F401
from .kkt import kkt_chol, kkt_chol2, kkt_ldl2, kkt_qr, KKTChol, cp_kktsolver, cpl_kktsolver  # noqa: F401
from . import scaling  # noqa: F401
from .conelp import conelp  # noqa: F401
from .batch import QPBatch, QPBatchGroup, qp_batch, qp_batch_distributed, shard_bounds, shard_indices  # noqa: F401

__all__ = ["kkt_chol", "kkt_chol2", "kkt_ldl2", "kkt_qr", "KKTChol", "cp_kktsolver", "cpl_kktsolver", "QPBatch", "qp_batch", "conelp", "qp_batch_distributed", "load",
           "device_count", "launch_count"]


def device_count():
    return load().cvxb_device_count()


def launch_count():
    return int(load().cvxb_launch_count())
