"""ctypes binding of the C++ host library (libcafe_host.so): the reference-named entry points
(load / tree / lambda / lambdamu / errormodel / pvalue / report commands, Nelder–Mead, prior, parsers)
that sit above the C-ABI.  Plumbing for tests and bench.py."""
from __future__ import annotations

import ctypes as C
from functools import lru_cache

import numpy as np

from . import buildlib as _build

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
MATH_FUNC = C.CFUNCTYPE(C.c_double, _dp, C.c_void_p)


class CafeHostError(RuntimeError):
    pass


@lru_cache(maxsize=None)
def load_library():
    _build.ensure_built()
    C.CDLL(_build.GPU_LIB, mode=C.RTLD_GLOBAL)
    H = C.CDLL(_build.HOST_LIB)
    vp = C.c_void_p
    H.cafe_host_last_error.restype = C.c_char_p
    H.cafe_host_gammaln.restype = C.c_double
    H.cafe_host_gammaln.argtypes = [C.c_double]
    H.cafe_host_chooseln.restype = C.c_double
    H.cafe_host_chooseln.argtypes = [C.c_double, C.c_double]
    H.cafe_host_poisspdf.restype = C.c_double
    H.cafe_host_poisspdf.argtypes = [C.c_int, C.c_double]
    H.cafe_host_pvalue.restype = C.c_double
    H.cafe_host_pvalue.argtypes = [C.c_double, _dp, C.c_int]
    H.cafe_host_lnc_table.argtypes = [C.c_int, _dp]
    H.cafe_host_init_family_size.argtypes = [C.c_int, _ip]
    H.cafe_host_fminsearch.argtypes = [MATH_FUNC, vp, C.c_int, _dp, C.c_double, C.c_double, _dp, _dp, _ip]
    H.cafe_host_parse_tree.argtypes = [C.c_char_p, C.c_int, _ip, _ip, _ip, _dp, C.c_char_p, C.c_int]
    H.cafe_host_parse_lambda_tree.argtypes = [C.c_char_p, C.c_char_p, _ip, C.c_int]
    H.cafe_host_read_errormodel.argtypes = [C.c_char_p, C.c_int, _dp, C.c_int, _ip, _ip]
    H.cafe_host_load_families.argtypes = [C.c_char_p, C.c_int, _ip, _ip, _ip, C.c_long, _ip, _ip]
    H.cafe_host_new.restype = vp
    H.cafe_host_new.argtypes = [C.c_int]
    H.cafe_host_free.argtypes = [vp]
    H.cafe_host_command.argtypes = [vp, C.c_char_p]
    H.cafe_host_num_params.argtypes = [vp]
    H.cafe_host_srand.argtypes = [C.c_uint]
    H.cafe_host_find_poisson_lambda.argtypes = [vp, _dp, _ip, _dp]
    H.cafe_host_get_family_table.argtypes = [vp, _ip, C.c_long, _ip, _ip]
    H.cafe_host_get_parameters.argtypes = [vp, _dp, C.c_int]
    H.cafe_host_objective_calls.argtypes = [vp]
    H.cafe_host_get_ranges.argtypes = [vp, _ip]
    H.cafe_host_get_prior.argtypes = [vp, _dp, C.c_int]
    H.cafe_host_num_families.argtypes = [vp]
    H.cafe_host_objective.argtypes = [vp, _dp, C.c_int, _dp]
    H.cafe_host_family_likelihoods.argtypes = [vp, _dp, C.c_long]
    H.cafe_host_get_cond_dist.argtypes = [vp, _dp, C.c_long, _ip, _ip]
    H.cafe_host_get_max_pvalues.argtypes = [vp, _dp, C.c_int]
    H.cafe_host_set_max_pvalues.argtypes = [vp, _dp, C.c_int]
    H.cafe_host_report_text_from.argtypes = [vp, _dp, C.c_int, _ip, _dp, _dp, _dp, C.c_char_p]
    H.cafe_host_chi2cdf.restype = C.c_double
    H.cafe_host_chi2cdf.argtypes = [C.c_double, C.c_int]
    H.cafe_host_likelihood_ratio_test.argtypes = [vp, C.c_int, _dp, C.c_long, _ip, _ip]
    H.cafe_host_branch_cutting.argtypes = [vp, C.c_int, C.c_int, _dp, C.c_long, _ip, _ip]
    return H


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


class FlatTree:
    """Tree in the reference's nlist (infix) order, as parsed by the C++ host."""

    def __init__(self, left, right, parent, branchlength, names):
        self.left = np.ascontiguousarray(left, dtype=np.int32)
        self.right = np.ascontiguousarray(right, dtype=np.int32)
        self.parent = np.ascontiguousarray(parent, dtype=np.int32)
        self.branchlength = np.ascontiguousarray(branchlength, dtype=np.float64)
        self.names = list(names)
        self.n_nodes = len(self.left)
        self.n_leaves = (self.n_nodes + 1) // 2
        self.root = int(np.where(self.parent < 0)[0][0])


def parse_tree(newick: str) -> FlatTree:
    H = load_library()
    cap = newick.count(",") * 2 + 8
    l = np.zeros(cap, dtype=np.int32)
    r = np.zeros(cap, dtype=np.int32)
    p = np.zeros(cap, dtype=np.int32)
    bl = np.zeros(cap)
    names = C.create_string_buffer(len(newick) + cap + 16)
    n = H.cafe_host_parse_tree(newick.encode(), cap, _i(l), _i(r), _i(p), _d(bl), names, len(names))
    if n < 0:
        raise CafeHostError(H.cafe_host_last_error().decode())
    nm = names.value.decode().split("\n")[:n]
    return FlatTree(l[:n], r[:n], p[:n], bl[:n], nm)


def parse_lambda_tree(tree_newick: str, lambda_newick: str):
    H = load_library()
    cap = tree_newick.count(",") * 2 + 8
    ids = np.zeros(cap, dtype=np.int32)
    m = H.cafe_host_parse_lambda_tree(tree_newick.encode(), lambda_newick.encode(), _i(ids), cap)
    if m < 0:
        raise CafeHostError(H.cafe_host_last_error().decode())
    n = tree_newick.count(",") * 2 + 1
    return m, ids[:n].copy()


def lnc_table(size: int):
    T = np.zeros((2 * size, size + 1))
    load_library().cafe_host_lnc_table(size, _d(T))
    return T


def init_family_size(mx: int):
    out = np.zeros(4, dtype=np.int32)
    load_library().cafe_host_init_family_size(mx, _i(out))
    return {"root_min": int(out[0]), "root_max": int(out[1]), "min": int(out[2]), "max": int(out[3])}


def prior_poisson(shift: int, lam: float, n: int = 1000):
    H = load_library()
    return np.array([H.cafe_host_poisspdf(shift - 1 + i, lam) for i in range(n)])


def read_errormodel(path: str, range_max: int):
    H = load_library()
    fd = C.c_int()
    td = C.c_int()
    dim = H.cafe_host_read_errormodel(path.encode(), range_max, None, 0, C.byref(fd), C.byref(td))
    if dim < 0:
        raise CafeHostError(H.cafe_host_last_error().decode())
    M = np.zeros((dim, dim))
    H.cafe_host_read_errormodel(path.encode(), range_max, _d(M), dim * dim, C.byref(fd), C.byref(td))
    return M, fd.value, td.value


def load_families(path: str, max_size: int = -1):
    H = load_library()
    ns = C.c_int()
    nf = C.c_int()
    ms = C.c_int()
    if H.cafe_host_load_families(path.encode(), max_size, C.byref(ns), C.byref(nf), None, 0, None, C.byref(ms)) != 0:
        raise CafeHostError(H.cafe_host_last_error().decode())
    counts = np.zeros((nf.value, ns.value), dtype=np.int32)
    ref = np.zeros(nf.value, dtype=np.int32)
    H.cafe_host_load_families(path.encode(), max_size, C.byref(ns), C.byref(nf), _i(counts), counts.size, _i(ref), C.byref(ms))
    return counts, ref, ms.value


def fminsearch(func, x0, tolx=1e-6, tolf=1e-6):
    """Run the host's Nelder–Mead on a Python callable f(list) -> float."""
    H = load_library()
    n = len(x0)

    def _cb(xp, _):
        return float(func([xp[i] for i in range(n)]))

    cb = MATH_FUNC(_cb)
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    xo = np.zeros(n)
    fo = C.c_double()
    it = C.c_int()
    rc = H.cafe_host_fminsearch(cb, None, n, _d(x0), tolx, tolf, _d(xo), C.byref(fo), C.byref(it))
    if rc != 0:
        raise CafeHostError(H.cafe_host_last_error().decode())
    return xo, fo.value, it.value


class Session:
    """The reference's Globals + command dispatcher (cafe_shell_dispatch_command)."""

    def __init__(self, quiet=True):
        self.H = load_library()
        self.h = self.H.cafe_host_new(1 if quiet else 0)

    def close(self):
        if self.h:
            self.H.cafe_host_free(self.h)
            self.h = None

    def command(self, line: str) -> int:
        return self.H.cafe_host_command(self.h, line.encode())

    def parameters(self):
        n = self.H.cafe_host_num_params(self.h)
        out = np.zeros(max(n, 1))
        self.H.cafe_host_get_parameters(self.h, _d(out), len(out))
        return out[:n]

    def ranges(self):
        out = np.zeros(4, dtype=np.int32)
        self.H.cafe_host_get_ranges(self.h, _i(out))
        return {"min": int(out[0]), "max": int(out[1]), "root_min": int(out[2]), "root_max": int(out[3])}

    def prior(self, n):
        out = np.zeros(n)
        if self.H.cafe_host_get_prior(self.h, _d(out), n) < 0:
            raise CafeHostError("prior not set")
        return out

    def find_poisson_lambda(self):
        lam = C.c_double()
        it = C.c_int()
        sc = C.c_double()
        if self.H.cafe_host_find_poisson_lambda(self.h, C.byref(lam), C.byref(it), C.byref(sc)) != 0:
            raise CafeHostError(self.H.cafe_host_last_error().decode())
        return lam.value, it.value, sc.value

    def family_table(self, n_species):
        F = self.num_families()
        counts = np.zeros((F, n_species), dtype=np.int32)
        ref = np.zeros(F, dtype=np.int32)
        index = np.zeros(n_species, dtype=np.int32)
        ns = self.H.cafe_host_get_family_table(self.h, _i(counts), counts.size, _i(ref), _i(index))
        if ns != n_species:
            raise CafeHostError("species count mismatch")
        return counts, ref, index

    def num_families(self):
        return self.H.cafe_host_num_families(self.h)

    def objective_calls(self):
        return self.H.cafe_host_objective_calls(self.h)

    def objective(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        out = C.c_double()
        if self.H.cafe_host_objective(self.h, _d(x), len(x), C.byref(out)) != 0:
            raise CafeHostError(self.H.cafe_host_last_error().decode())
        return out.value

    def family_likelihoods(self):
        F = self.num_families()
        rg = self.ranges()
        R = rg["root_max"] - rg["root_min"] + 1
        out = np.zeros((F, R))
        if self.H.cafe_host_family_likelihoods(self.h, _d(out), out.size) < 0:
            raise CafeHostError(self.H.cafe_host_last_error().decode())
        return out

    def cond_dist(self):
        rows = C.c_int()
        cols = C.c_int()
        buf = np.zeros(1000 * 20000)
        if self.H.cafe_host_get_cond_dist(self.h, _d(buf), buf.size, C.byref(rows), C.byref(cols)) != 0:
            raise CafeHostError("no conditional distribution")
        return buf[: rows.value * cols.value].reshape(rows.value, cols.value).copy()

    def report_text_from(self, lambdas, sizes, branch_pv, max_pv, path, likelihood_ratios=None):
        """Write the text report (cafe_report_text) from given per-family results; no device work."""
        lam = np.ascontiguousarray(lambdas, dtype=np.float64)
        sizes = np.ascontiguousarray(sizes, dtype=np.int32)
        bpv = np.ascontiguousarray(branch_pv, dtype=np.float64)
        mpv = np.ascontiguousarray(max_pv, dtype=np.float64)
        lr = None if likelihood_ratios is None else np.ascontiguousarray(likelihood_ratios, dtype=np.float64)
        rc = self.H.cafe_host_report_text_from(self.h, _d(lam), len(lam), _i(sizes), _d(bpv), _d(mpv),
                                               None if lr is None else _d(lr), str(path).encode())
        if rc < 0:
            raise CafeHostError(self.H.cafe_host_last_error().decode())

    def set_max_pvalues(self, pv):
        pv = np.ascontiguousarray(pv, dtype=np.float64)
        self.H.cafe_host_set_max_pvalues(self.h, _d(pv), len(pv))

    def likelihood_ratio_test(self, tree_level_mu=False):
        """cafe_likelihood_ratio_test (cafe/cafe_main.c:398-431): likelihoodRatios [nodes][families].  tree_level_mu=True keys the
        lengthened branches with mu = 0 like the stock reference binary (its tree copy drops the nodes' mu)."""
        F = self.num_families()
        buf = np.zeros(F * 4096)
        nodes = C.c_int()
        fams = C.c_int()
        if self.H.cafe_host_likelihood_ratio_test(self.h, int(tree_level_mu), _d(buf), buf.size, C.byref(nodes), C.byref(fams)) < 0:
            raise CafeHostError(self.H.cafe_host_last_error().decode())
        return buf[: nodes.value * fams.value].reshape(nodes.value, fams.value).copy()

    def branch_cutting(self, num_random_samples, tree_level_mu=False):
        """cafe_branch_cutting (cafe/branch_cutting.cpp:101-272): cutPvalues [nodes][families] for the family p-values given with
        set_max_pvalues (or the last report).  The conditional distributions replay glibc rand() (after `seed`) when the session
        has one thread, as the reference's single-threaded run would draw them."""
        F = self.num_families()
        buf = np.zeros(F * 4096)
        nodes = C.c_int()
        fams = C.c_int()
        if self.H.cafe_host_branch_cutting(self.h, int(num_random_samples), int(tree_level_mu), _d(buf), buf.size, C.byref(nodes), C.byref(fams)) < 0:
            raise CafeHostError(self.H.cafe_host_last_error().decode())
        return buf[: nodes.value * fams.value].reshape(nodes.value, fams.value).copy()

    def max_pvalues(self):
        F = self.num_families()
        out = np.zeros(F)
        n = self.H.cafe_host_get_max_pvalues(self.h, _d(out), F)
        return out[:n]


def chi2cdf(x: float, df: int = 1) -> float:
    return load_library().cafe_host_chi2cdf(float(x), int(df))


def srand(seed: int):
    load_library().cafe_host_srand(seed)


def release_gpu():
    load_library().cafe_host_release_gpu()
