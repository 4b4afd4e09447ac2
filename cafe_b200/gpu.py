"""ctypes binding of the C-ABI in include/cafe_gpu.h (libcafe_gpu.so).

This is plumbing for tests and bench.py: numpy arrays in, numpy arrays out, every call goes through
the C-ABI entry points a reference-side binding would use.  There is no fallback: if the library
or a CUDA device is missing, construction raises.
"""
from __future__ import annotations

import ctypes as C
from functools import lru_cache

import numpy as np

from . import buildlib as _build

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)

ZERO_LIKELIHOOD = 1

# every symbol include/cafe_gpu.h declares (checked by tests/test_abi.py)
ABI_SYMBOLS = [
    "cafe_gpu_abi_version", "cafe_gpu_create", "cafe_gpu_destroy", "cafe_gpu_last_error", "cafe_gpu_set_stream",
    "cafe_gpu_synchronize", "cafe_gpu_set_tree", "cafe_gpu_set_ranges", "cafe_gpu_set_lnc_table",
    "cafe_gpu_set_families", "cafe_gpu_set_prior", "cafe_gpu_set_error_model", "cafe_gpu_set_rates",
    "cafe_gpu_build_matrices", "cafe_gpu_num_keys", "cafe_gpu_get_matrix", "cafe_gpu_score", "cafe_gpu_objective",
    "cafe_gpu_objective_device", "cafe_gpu_family_results", "cafe_gpu_family_likelihoods",
    "cafe_gpu_conditional_distribution", "cafe_gpu_pvalues", "cafe_gpu_cut_pvalues", "cafe_gpu_launch_count",
    "cafe_gpu_reset_launch_count", "cafe_gpu_enable_timing", "cafe_gpu_timing_collect", "cafe_gpu_score_flops",
    "cafe_gpu_score_device", "cafe_gpu_set_key_shard", "cafe_gpu_matrix_storage", "cafe_gpu_matrices_exchanged",
    "cafe_gpu_viterbi", "cafe_gpu_viterbi_report", "cafe_gpu_conditional_distribution_rows",
    "cafe_gpu_likelihood_ratio_test", "cafe_gpu_timing_collect4", "cafe_gpu_comm_unique_id", "cafe_gpu_comm_init",
    "cafe_gpu_comm_size", "cafe_gpu_comm_rank", "cafe_gpu_create_multi", "cafe_gpu_num_devices",
]

COMM_ID_BYTES = 128


class CafeGpuError(RuntimeError):
    pass


@lru_cache(maxsize=None)
def load_library():
    _build.ensure_built()
    L = C.CDLL(_build.GPU_LIB)
    vp = C.c_void_p
    L.cafe_gpu_abi_version.restype = C.c_int
    L.cafe_gpu_create.argtypes = [C.POINTER(vp), C.c_int]
    L.cafe_gpu_destroy.argtypes = [vp]
    L.cafe_gpu_last_error.restype = C.c_char_p
    L.cafe_gpu_last_error.argtypes = [vp]
    L.cafe_gpu_set_stream.argtypes = [vp, vp]
    L.cafe_gpu_synchronize.argtypes = [vp]
    L.cafe_gpu_set_tree.argtypes = [vp, C.c_int, _ip, _ip, _dp]
    L.cafe_gpu_set_ranges.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int]
    L.cafe_gpu_set_lnc_table.argtypes = [vp, _dp, C.c_int, C.c_int]
    L.cafe_gpu_set_families.argtypes = [vp, C.c_int, C.c_int, _ip, _ip, _ip]
    L.cafe_gpu_set_prior.argtypes = [vp, _dp, C.c_int]
    L.cafe_gpu_set_error_model.argtypes = [vp, C.c_int, _dp, C.c_int]
    L.cafe_gpu_set_rates.argtypes = [vp, _dp, _dp]
    L.cafe_gpu_build_matrices.argtypes = [vp]
    L.cafe_gpu_num_keys.argtypes = [vp]
    L.cafe_gpu_get_matrix.argtypes = [vp, C.c_int, _dp, C.c_int]
    L.cafe_gpu_score.argtypes = [vp, _dp, _ip]
    L.cafe_gpu_objective.argtypes = [vp, _dp, _dp, _dp, _ip]
    L.cafe_gpu_objective_device.argtypes = [vp, _dp, _dp, vp]
    L.cafe_gpu_score_device.argtypes = [vp, vp]
    L.cafe_gpu_set_key_shard.argtypes = [vp, C.c_int, C.c_int]
    L.cafe_gpu_matrix_storage.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(C.c_int64), C.POINTER(C.c_int32)]
    L.cafe_gpu_matrices_exchanged.argtypes = [vp]
    L.cafe_gpu_viterbi.argtypes = [vp, _ip, _dp]
    L.cafe_gpu_viterbi_report.argtypes = [vp, _ip, _dp]
    L.cafe_gpu_family_results.argtypes = [vp, _dp, _dp, _ip]
    L.cafe_gpu_likelihood_ratio_test.argtypes = [vp, C.POINTER(C.c_uint8), _dp, _dp, _dp, _ip]
    L.cafe_gpu_family_likelihoods.argtypes = [vp, _dp]
    L.cafe_gpu_conditional_distribution.argtypes = [vp, C.c_int, _dp, C.c_uint64, _dp]
    L.cafe_gpu_conditional_distribution_rows.argtypes = [vp, C.c_int, _dp, C.c_uint64, C.c_int, C.c_int, _dp]
    L.cafe_gpu_pvalues.argtypes = [vp, _dp, C.c_int, C.c_int, _dp]
    L.cafe_gpu_cut_pvalues.argtypes = [vp, _dp, _dp, C.c_int, C.c_int, _dp, _dp, C.c_int, _dp]
    L.cafe_gpu_launch_count.restype = C.c_int64
    L.cafe_gpu_launch_count.argtypes = [vp]
    L.cafe_gpu_reset_launch_count.argtypes = [vp]
    L.cafe_gpu_enable_timing.argtypes = [vp, C.c_int]
    L.cafe_gpu_timing_collect.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int]
    L.cafe_gpu_timing_collect4.argtypes = [vp] + [C.POINTER(C.c_float)] * 4 + [C.c_int]
    L.cafe_gpu_comm_unique_id.argtypes = [vp, C.c_int]
    L.cafe_gpu_comm_init.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int]
    L.cafe_gpu_comm_size.argtypes = [vp]
    L.cafe_gpu_comm_rank.argtypes = [vp]
    L.cafe_gpu_create_multi.argtypes = [C.POINTER(vp), _ip, C.c_int]
    L.cafe_gpu_num_devices.argtypes = [vp]
    L.cafe_gpu_score_flops.restype = C.c_double
    L.cafe_gpu_score_flops.argtypes = [vp]
    return L


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


def comm_unique_id() -> bytes:
    """ncclGetUniqueId through the C-ABI: call on one rank, hand the bytes to every rank (cafe_gpu_comm_init)."""
    L = load_library()
    buf = C.create_string_buffer(COMM_ID_BYTES)
    rc = L.cafe_gpu_comm_unique_id(buf, COMM_ID_BYTES)
    if rc != 0:
        raise CafeGpuError(f"cafe_gpu_comm_unique_id failed ({rc})")
    return buf.raw


class CafeGpu:
    """One C-ABI context: one device and one stream, or (devices=[...]) the leader of one context per device in this
    process (cafe_gpu_create_multi), which shards the families and runs the NCCL exchange steps inside the library."""

    def __init__(self, device: int = -1, devices=None):
        self.L = load_library()
        h = C.c_void_p()
        if devices is not None:
            dv = np.ascontiguousarray(devices, dtype=np.int32)
            rc = self.L.cafe_gpu_create_multi(C.byref(h), _i(dv), len(dv))
        else:
            rc = self.L.cafe_gpu_create(C.byref(h), device)
        if rc != 0:
            raise CafeGpuError(f"cafe_gpu_create failed ({rc}): {self.L.cafe_gpu_last_error(None).decode()}")
        self.h = h
        self.n_nodes = 0
        self.n_leaves = 0
        self.R = 0
        self.S = 0
        self.F = 0

    def close(self):
        if getattr(self, "h", None):
            self.L.cafe_gpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, what):
        if rc < 0:
            raise CafeGpuError(f"{what} failed ({rc}): {self.L.cafe_gpu_last_error(self.h).decode()}")
        return rc

    def set_stream(self, cuda_stream_ptr):
        self._ck(self.L.cafe_gpu_set_stream(self.h, C.c_void_p(cuda_stream_ptr)), "set_stream")

    def comm_init(self, unique_id: bytes, rank: int, world: int):
        """Make this context rank `rank` of an NCCL communicator of `world` contexts (one process per GPU).  From then on
        build_matrices shards K1 over the ranks and score/objective return the all-reduced result on every rank."""
        buf = C.create_string_buffer(unique_id, COMM_ID_BYTES)
        self._ck(self.L.cafe_gpu_comm_init(self.h, buf, COMM_ID_BYTES, rank, world), "comm_init")

    def num_devices(self):
        return self.L.cafe_gpu_num_devices(self.h)

    def synchronize(self):
        self._ck(self.L.cafe_gpu_synchronize(self.h), "synchronize")

    def set_tree(self, left, right, branchlength):
        left = np.ascontiguousarray(left, dtype=np.int32)
        right = np.ascontiguousarray(right, dtype=np.int32)
        bl = np.ascontiguousarray(branchlength, dtype=np.float64)
        self._ck(self.L.cafe_gpu_set_tree(self.h, len(left), _i(left), _i(right), _d(bl)), "set_tree")
        self.n_nodes = len(left)
        self.n_leaves = (len(left) + 1) // 2

    def set_ranges(self, rmin, rmax, root_min, root_max):
        self._ck(self.L.cafe_gpu_set_ranges(self.h, rmin, rmax, root_min, root_max), "set_ranges")
        self.R = root_max - root_min + 1
        self.S = max(rmax, root_max) + 1

    def set_lnc_table(self, table):
        table = np.ascontiguousarray(table, dtype=np.float64)
        self._ck(self.L.cafe_gpu_set_lnc_table(self.h, _d(table), table.shape[0], table.shape[1]), "set_lnc_table")

    def set_families(self, counts, multiplicity=None, first_index=None):
        counts = np.ascontiguousarray(counts, dtype=np.int32)
        m = None if multiplicity is None else np.ascontiguousarray(multiplicity, dtype=np.int32)
        fi = None if first_index is None else np.ascontiguousarray(first_index, dtype=np.int32)
        self._ck(self.L.cafe_gpu_set_families(self.h, counts.shape[0], counts.shape[1], _i(counts),
                                              None if m is None else _i(m), None if fi is None else _i(fi)),
                 "set_families")
        self.F = counts.shape[0]

    def set_prior(self, prior):
        prior = np.ascontiguousarray(prior, dtype=np.float64)
        self._ck(self.L.cafe_gpu_set_prior(self.h, _d(prior), len(prior)), "set_prior")

    def set_error_model(self, leaf, matrix):
        if matrix is None:
            self._ck(self.L.cafe_gpu_set_error_model(self.h, leaf, None, 0), "set_error_model")
            return
        matrix = np.ascontiguousarray(matrix, dtype=np.float64)
        self._ck(self.L.cafe_gpu_set_error_model(self.h, leaf, _d(matrix), matrix.shape[0]), "set_error_model")

    def set_rates(self, lam, mu):
        lam = np.ascontiguousarray(lam, dtype=np.float64)
        mu = np.ascontiguousarray(mu, dtype=np.float64)
        self._ck(self.L.cafe_gpu_set_rates(self.h, _d(lam), _d(mu)), "set_rates")

    def build_matrices(self):
        self._ck(self.L.cafe_gpu_build_matrices(self.h), "build_matrices")

    def num_keys(self):
        return self.L.cafe_gpu_num_keys(self.h)

    def get_matrix(self, node):
        out = np.zeros((self.S, self.S))
        self._ck(self.L.cafe_gpu_get_matrix(self.h, node, _d(out), self.S), "get_matrix")
        return out

    def score(self):
        s = C.c_double()
        fz = C.c_int32(-1)
        rc = self._ck(self.L.cafe_gpu_score(self.h, C.byref(s), C.byref(fz)), "score")
        return s.value, (fz.value if rc == ZERO_LIKELIHOOD else -1)

    def objective(self, lam, mu):
        lam = np.ascontiguousarray(lam, dtype=np.float64)
        mu = np.ascontiguousarray(mu, dtype=np.float64)
        s = C.c_double()
        fz = C.c_int32(-1)
        rc = self._ck(self.L.cafe_gpu_objective(self.h, _d(lam), _d(mu), C.byref(s), C.byref(fz)), "objective")
        return s.value, (fz.value if rc == ZERO_LIKELIHOOD else -1)

    def objective_device(self, lam, mu, out_device_ptr):
        lam = np.ascontiguousarray(lam, dtype=np.float64)
        mu = np.ascontiguousarray(mu, dtype=np.float64)
        self._ck(self.L.cafe_gpu_objective_device(self.h, _d(lam), _d(mu), C.c_void_p(out_device_ptr)), "objective_device")

    def score_device(self, out_device_ptr):
        self._ck(self.L.cafe_gpu_score_device(self.h, C.c_void_p(out_device_ptr)), "score_device")

    # ---- K1 sharded across ranks (include/cafe_gpu.h: cafe_gpu_set_key_shard) ----
    def set_key_shard(self, rank, world):
        self._ck(self.L.cafe_gpu_set_key_shard(self.h, rank, world), "set_key_shard")

    def matrix_storage(self):
        """(d_M, d_MT, doubles_per_key, keys_per_rank): device pointers of the two matrix buffers."""
        pm, pt = C.c_void_p(), C.c_void_p()
        dpk, kpr = C.c_int64(), C.c_int32()
        self._ck(self.L.cafe_gpu_matrix_storage(self.h, C.byref(pm), C.byref(pt), C.byref(dpk), C.byref(kpr)), "matrix_storage")
        return pm.value, pt.value, dpk.value, kpr.value

    def matrices_exchanged(self):
        self._ck(self.L.cafe_gpu_matrices_exchanged(self.h), "matrices_exchanged")

    def family_results(self):
        lp = np.zeros(self.F)
        ml = np.zeros(self.F)
        am = np.zeros(self.F, dtype=np.int32)
        self._ck(self.L.cafe_gpu_family_results(self.h, _d(lp), _d(ml), _i(am)), "family_results")
        return lp, ml, am

    def viterbi(self):
        """(sizes[F][n_nodes] int32 in nlist order, max root likelihood[F]) — cafe_tree_viterbi for every family."""
        sizes = np.zeros((self.F, self.n_nodes), dtype=np.int32)
        ml = np.zeros(self.F)
        self._ck(self.L.cafe_gpu_viterbi(self.h, _i(sizes), _d(ml)), "viterbi")
        return sizes, ml

    def viterbi_report(self):
        """(sizes[F][n_nodes], branch p-values[F][n_nodes]) with every family's forced range, as viterbi_section does."""
        sizes = np.zeros((self.F, self.n_nodes), dtype=np.int32)
        pv = np.zeros((self.F, self.n_nodes))
        self._ck(self.L.cafe_gpu_viterbi_report(self.h, _i(sizes), _d(pv)), "viterbi_report")
        return sizes, pv

    def likelihood_ratio_test(self, tested=None, lengthened_mu=None):
        """(base max likelihood [F], best max likelihood [n_nodes][F], steps [n_nodes][F]) of the branch-stretch test."""
        base = np.zeros(self.F)
        best = np.zeros((self.n_nodes, self.F))
        steps = np.zeros((self.n_nodes, self.F), dtype=np.int32)
        tp = None
        if tested is not None:
            tested = np.ascontiguousarray(tested, dtype=np.uint8)
            assert tested.shape == (self.F,)
            tp = tested.ctypes.data_as(C.POINTER(C.c_uint8))
        mp = None
        if lengthened_mu is not None:
            lengthened_mu = np.ascontiguousarray(lengthened_mu, dtype=np.float64)
            assert lengthened_mu.shape == (self.n_nodes,)
            mp = _d(lengthened_mu)
        self._ck(self.L.cafe_gpu_likelihood_ratio_test(self.h, tp, mp, _d(base), _d(best), _i(steps)), "likelihood_ratio_test")
        return base, best, steps

    def family_likelihoods(self):
        out = np.zeros((self.F, self.R))
        self._ck(self.L.cafe_gpu_family_likelihoods(self.h, _d(out)), "family_likelihoods")
        return out

    def conditional_distribution(self, n_samples, uniforms=None, seed=0):
        out = np.zeros((self.R, n_samples))
        up = None
        if uniforms is not None:
            uniforms = np.ascontiguousarray(uniforms, dtype=np.float64)
            need = self.R * n_samples * (self.n_nodes - 1)
            if uniforms.size < need:
                raise ValueError(f"replay stream needs {need} uniforms")
            up = _d(uniforms)
        self._ck(self.L.cafe_gpu_conditional_distribution(self.h, n_samples, up, C.c_uint64(seed), _d(out)),
                 "conditional_distribution")
        return out

    def conditional_distribution_rows(self, n_samples, row_lo, row_hi, seed=0):
        out = np.zeros((row_hi - row_lo, n_samples))
        self._ck(self.L.cafe_gpu_conditional_distribution_rows(self.h, n_samples, None, C.c_uint64(seed), row_lo, row_hi, _d(out)),
                 "conditional_distribution_rows")
        return out

    def pvalues(self, cd):
        cd = np.ascontiguousarray(cd, dtype=np.float64)
        out = np.zeros(self.F)
        self._ck(self.L.cafe_gpu_pvalues(self.h, _d(cd), cd.shape[0], cd.shape[1], _d(out)), "pvalues")
        return out

    def cut_pvalues(self, L_rest, cd_rest, L_sub=None, cd_sub=None):
        """Branch cutting, the p-value part: rows [F][rfsize] of one or both sides of the cut with their distributions."""
        L_rest = np.ascontiguousarray(L_rest, dtype=np.float64); cd_rest = np.ascontiguousarray(cd_rest, dtype=np.float64)
        F, rf = L_rest.shape
        out = np.zeros(F)
        l2 = c2 = None
        if L_sub is not None:
            L_sub = np.ascontiguousarray(L_sub, dtype=np.float64); cd_sub = np.ascontiguousarray(cd_sub, dtype=np.float64)
            l2, c2 = _d(L_sub), _d(cd_sub)
        self._ck(self.L.cafe_gpu_cut_pvalues(self.h, _d(L_rest), l2, F, rf, _d(cd_rest), c2, cd_rest.shape[1], _d(out)), "cut_pvalues")
        return out

    def launch_count(self):
        return int(self.L.cafe_gpu_launch_count(self.h))

    def reset_launch_count(self):
        self.L.cafe_gpu_reset_launch_count(self.h)

    def enable_timing(self, on=True):
        self._ck(self.L.cafe_gpu_enable_timing(self.h, 1 if on else 0), "enable_timing")

    def timing_collect(self, cap=256):
        k1 = np.zeros(cap, dtype=np.float32)
        k2 = np.zeros(cap, dtype=np.float32)
        fp = C.POINTER(C.c_float)
        n = self._ck(self.L.cafe_gpu_timing_collect(self.h, k1.ctypes.data_as(fp), k2.ctypes.data_as(fp), cap), "timing_collect")
        return k1[:n].copy(), k2[:n].copy()

    def timing_collect4(self, cap=256):
        """(K1, matrix exchange, K2, score reduction) device times in ms per evaluation."""
        a = [np.zeros(cap, dtype=np.float32) for _ in range(4)]
        fp = C.POINTER(C.c_float)
        n = self._ck(self.L.cafe_gpu_timing_collect4(self.h, *[x.ctypes.data_as(fp) for x in a], cap), "timing_collect4")
        return tuple(x[:n].copy() for x in a)

    def score_flops(self):
        return float(self.L.cafe_gpu_score_flops(self.h))
