"""ctypes loaders for the CPU oracle — TEST INFRASTRUCTURE ONLY.

`oracle.lib()`   -> oracle/libcafe_oracle.so  (our C restatement, oracle/cafe_oracle.c)
`oracle.ref()`   -> oracle/_ref/libcafe_ref.so (the unmodified reference + oracle/ref_shim.cpp), or None
                    when it has not been built (it needs /root/reference at build time; the built .so
                    travels to the GPU box, the sources do not).

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may import this
package.  The product (cafe_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from functools import lru_cache

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_lp = C.POINTER(C.c_long)
_dpp = C.POINTER(_dp)
MATH_FUNC = C.CFUNCTYPE(C.c_double, _dp, C.c_void_p)


def build(ref: bool = True) -> None:
    """Compile the oracle (always) and the reference (when /root/reference is present)."""
    subprocess.run(["make", "-s", "-C", HERE, "oracle"], check=True)
    if ref and os.path.isdir("/root/reference/cafe"):
        subprocess.run(["make", "-s", "-j8", "-C", HERE, "ref"], check=True)
        # the reference with integration/gpu_bridge.cpp compiled in (needs ../cafe_b200/libcafe_gpu.so)
        if os.path.exists(os.path.join(HERE, "..", "cafe_b200", "libcafe_gpu.so")):
            subprocess.run(["make", "-s", "-j8", "-C", HERE, "bridge"], check=True)


def _dptr(a):
    return a.ctypes.data_as(_dp)


def _iptr(a):
    return a.ctypes.data_as(_ip)


@lru_cache(maxsize=None)
def lib():
    path = os.path.join(HERE, "libcafe_oracle.so")
    if not os.path.exists(path):
        build(ref=False)
    L = C.CDLL(path)
    L.orc_gammaln.restype = C.c_double
    L.orc_gammaln.argtypes = [C.c_double]
    L.orc_chooseln.restype = C.c_double
    L.orc_chooseln.argtypes = [C.c_double, C.c_double]
    L.orc_poisspdf.restype = C.c_double
    L.orc_poisspdf.argtypes = [C.c_int, C.c_double]
    L.orc_unifrnd.restype = C.c_double
    L.orc_lnc_table.argtypes = [C.c_int, _dp]
    L.orc_bd_matrix.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int, _dp]
    L.orc_key_branchlength.restype = C.c_int
    L.orc_key_branchlength.argtypes = [C.c_double]
    L.orc_matvec.argtypes = [_dp, C.c_int, _dp, C.c_int, C.c_int, C.c_int, C.c_int, _dp]
    L.orc_prune.restype = C.c_int
    L.orc_prune.argtypes = [C.c_int, _ip, _ip, C.c_int, _dpp, C.c_int, _ip, _dpp, C.c_int,
                            C.c_int, C.c_int, C.c_int, C.c_int, _dp]
    L.orc_posterior.argtypes = [_dp, _dp, C.c_int, _dp, _dp, _ip]
    L.orc_viterbi_branch_pvalues.argtypes = [C.c_int, _ip, _ip, C.c_int, _dpp, C.c_int, _ip, C.c_int, _dp]
    L.orc_viterbi.argtypes = [C.c_int, _ip, _ip, C.c_int, _dpp, C.c_int, _ip, _dpp, C.c_int,
                              C.c_int, C.c_int, C.c_int, C.c_int, _ip, _dp]
    L.orc_score.restype = C.c_double
    L.orc_score.argtypes = [C.c_int, _ip, _ip, C.c_int, _dpp, C.c_int, _dpp, C.c_int,
                            C.c_int, C.c_int, C.c_int, C.c_int, _dp, C.c_int, _ip, _ip,
                            _dp, _dp, _ip, _dp, _ip]
    L.orc_random_familysize.restype = C.c_int
    L.orc_random_familysize.argtypes = [C.c_int, _ip, _ip, C.c_int, _dpp, C.c_int, C.c_int, C.c_int, _dp, _lp, _ip]
    L.orc_random_probabilities.argtypes = [C.c_int, _ip, _ip, C.c_int, _dpp, C.c_int, C.c_int, C.c_int,
                                           C.c_int, C.c_int, _dp, _lp, _dp, _dp, _ip, _ip, _dpp, C.c_int]
    L.orc_pvalue.restype = C.c_double
    L.orc_pvalue.argtypes = [C.c_double, _dp, C.c_int]
    L.orc_init_family_size.argtypes = [C.c_int, _ip, _ip, _ip, _ip]
    L.orc_family_pvalue.restype = C.c_double
    L.orc_family_pvalue.argtypes = [C.c_int, _ip, _ip, C.c_int, _dpp, C.c_int, _dpp, C.c_int, _ip,
                                    _dp, C.c_int, C.c_int, _dp, _ip]
    L.orc_chi2cdf.restype = C.c_double
    L.orc_chi2cdf.argtypes = [C.c_double, C.c_int]
    L.orc_lrt_family.restype = C.c_int
    L.orc_lrt_family.argtypes = [C.c_int, _ip, _ip, C.c_int, _dpp, C.c_int, _dp, _dp, _dp, _ip, _dpp, C.c_int,
                                 C.c_int, C.c_int, C.c_int, C.c_int, _dp, _dp, _ip]
    return L


@lru_cache(maxsize=None)
def ref(openmp: bool = False):
    """The compiled, unmodified reference (oracle/_ref, built by oracle/Makefile) or None.  openmp=True: the build with
    -fopenmp, as the reference's own Makefile.in:13 compiles it (bench.py's reference arm)."""
    path = os.path.join(HERE, "_ref", "libcafe_ref_omp.so" if openmp else "libcafe_ref.so")
    if not os.path.exists(path):
        return None
    R = C.CDLL(path)
    R.refshim_gammaln.restype = C.c_double
    R.refshim_gammaln.argtypes = [C.c_double]
    R.refshim_chooseln.restype = C.c_double
    R.refshim_chooseln.argtypes = [C.c_double, C.c_double]
    R.refshim_poisspdf.restype = C.c_double
    R.refshim_poisspdf.argtypes = [C.c_int, C.c_double]
    R.refshim_pvalue.restype = C.c_double
    R.refshim_pvalue.argtypes = [C.c_double, _dp, C.c_int]
    R.refshim_srand.argtypes = [C.c_uint]
    R.refshim_unifrnd.restype = C.c_double
    R.refshim_init_family_size.argtypes = [C.c_int, _ip]
    R.refshim_bd_matrix.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int, _dp]
    R.refshim_bd_rate_log_alpha.restype = C.c_double
    R.refshim_bd_rate_log_alpha.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.c_int]
    R.refshim_bd_likelihood_s_c.restype = C.c_double
    R.refshim_bd_likelihood_s_c.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int]
    R.refshim_matvec.argtypes = [_dp, C.c_int, _dp, C.c_int, C.c_int, C.c_int, C.c_int, _dp]
    R.refshim_session_new.restype = C.c_void_p
    R.refshim_session_new.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int]
    R.refshim_n_nodes.restype = C.c_int
    R.refshim_n_nodes.argtypes = [C.c_void_p]
    R.refshim_describe.argtypes = [C.c_void_p, _ip, _ip, _ip, _dp, C.c_char_p, C.c_int]
    R.refshim_set_range.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
    R.refshim_set_rates.argtypes = [C.c_void_p, _dp, _dp]
    R.refshim_reset_cache.argtypes = [C.c_void_p]
    R.refshim_get_matrix.restype = C.c_int
    R.refshim_get_matrix.argtypes = [C.c_void_p, C.c_int, _dp]
    R.refshim_set_errormodel.argtypes = [C.c_void_p, C.c_int, _dp, C.c_int, C.c_int, C.c_int]
    R.refshim_read_errormodel.restype = C.c_int
    R.refshim_read_errormodel.argtypes = [C.c_char_p, C.c_int, _dp, _ip, _ip]
    R.refshim_viterbi_forced.argtypes = [C.c_void_p, C.c_int, _ip, _dp]
    R.refshim_viterbi.restype = C.c_double
    R.refshim_viterbi.argtypes = [C.c_void_p, _ip, _ip]
    R.refshim_likelihoods.restype = C.c_int
    R.refshim_likelihoods.argtypes = [C.c_void_p, _ip, _dp]
    R.refshim_set_families.argtypes = [C.c_void_p, C.c_int, _ip, C.c_int]
    R.refshim_get_refs.argtypes = [C.c_void_p, _ip]
    R.refshim_get_posterior.restype = C.c_double
    R.refshim_get_posterior.argtypes = [C.c_void_p, _dp, _ip, C.c_char_p, C.c_int]
    R.refshim_get_maxlh.argtypes = [C.c_void_p, _ip]
    R.refshim_prior_poisson.argtypes = [C.c_int, C.c_double, _dp]
    R.refshim_find_poisson_lambda.restype = C.c_double
    R.refshim_find_poisson_lambda.argtypes = [C.c_void_p, _ip, _dp]
    R.refshim_cond_dist.restype = C.c_int
    R.refshim_cond_dist.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp]
    R.refshim_random_probabilities.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp]
    R.refshim_random_familysize.restype = C.c_int
    R.refshim_random_familysize.argtypes = [C.c_void_p, C.c_int, C.c_int, _ip]
    R.refshim_family_pvalue.restype = C.c_double
    R.refshim_family_pvalue.argtypes = [C.c_void_p, C.c_int, _dp, C.c_int, C.c_int, _dp, _ip]
    R.refshim_load_families.restype = C.c_int
    R.refshim_load_families.argtypes = [C.c_char_p, C.c_int, _ip, _ip, _ip, C.c_int, _ip, _ip]
    R.refshim_session_free.argtypes = [C.c_void_p]
    R.refshim_chi2cdf.restype = C.c_double
    R.refshim_chi2cdf.argtypes = [C.c_double, C.c_int]
    R.refshim_likelihood_ratio_test.argtypes = [C.c_void_p, _dp, C.c_double, C.c_int, _dp]
    R.refshim_fminsearch.argtypes = [MATH_FUNC, C.c_void_p, C.c_int, _dp, C.c_double, C.c_double, _dp, _dp, _ip]
    return R


def ref_gpu_binary():
    """The unmodified reference + integration/gpu_bridge.cpp (oracle/Makefile target `bridge`), or None."""
    p = os.path.join(HERE, "_ref", "cafe_ref_gpu")
    return p if os.path.exists(p) else None


def ref_binary():
    p = os.path.join(HERE, "_ref", "cafe_ref")
    return p if os.path.exists(p) else None


# --------------------------------------------------------------------------- flat trees

class FlatTree:
    """Binary tree in the reference's nlist (infix) numbering: leaves even, internal odd
    (cafe/cafe_commands.cpp:1985-2051)."""

    def __init__(self, left, right, parent, branchlength, names):
        self.left = np.ascontiguousarray(left, dtype=np.int32)
        self.right = np.ascontiguousarray(right, dtype=np.int32)
        self.parent = np.ascontiguousarray(parent, dtype=np.int32)
        self.branchlength = np.ascontiguousarray(branchlength, dtype=np.float64)
        self.names = list(names)
        self.n_nodes = len(self.left)
        self.n_leaves = (self.n_nodes + 1) // 2
        self.root = int(np.where(self.parent < 0)[0][0])

    @property
    def leaf_names(self):
        return [self.names[i] for i in range(0, self.n_nodes, 2)]


def parse_newick(s: str) -> FlatTree:
    """Minimal Newick reader (names + ':length', binary only) producing the nlist numbering.
    Test-side convenience; the product's parser is cafe_b200/host (C++)."""
    s = s.strip().rstrip(";")
    pos = 0

    def node():
        nonlocal pos
        kids = []
        if s[pos] == "(":
            pos += 1
            while True:
                kids.append(node())
                if s[pos] == ",":
                    pos += 1
                    continue
                assert s[pos] == ")", f"bad newick at {pos}"
                pos += 1
                break
        start = pos
        while pos < len(s) and s[pos] not in ",():;":
            pos += 1
        name = s[start:pos]
        bl = -1.0
        if pos < len(s) and s[pos] == ":":
            pos += 1
            start = pos
            while pos < len(s) and s[pos] not in ",()":
                pos += 1
            bl = float(s[start:pos])
        return {"kids": kids, "name": name, "bl": bl}

    rootn = node()
    order = []

    def infix(n):
        if n["kids"]:
            assert len(n["kids"]) == 2, "Tree must be binary"
            infix(n["kids"][0])
            order.append(n)
            infix(n["kids"][1])
        else:
            order.append(n)

    infix(rootn)
    for i, n in enumerate(order):
        n["id"] = i
    N = len(order)
    left = [-1] * N
    right = [-1] * N
    parent = [-1] * N
    for n in order:
        if n["kids"]:
            left[n["id"]] = n["kids"][0]["id"]
            right[n["id"]] = n["kids"][1]["id"]
            parent[n["kids"][0]["id"]] = n["id"]
            parent[n["kids"][1]["id"]] = n["id"]
    return FlatTree(left, right, parent, [n["bl"] for n in order], [n["name"] for n in order])


# --------------------------------------------------------------------------- numpy-level helpers

def bd_matrix(t, lam, mu, maxfs):
    M = np.zeros((maxfs + 1, maxfs + 1))
    lib().orc_bd_matrix(float(t), float(lam), float(mu), int(maxfs), _dptr(M))
    return M


def lnc_table(size):
    T = np.zeros((2 * size, size + 1))
    lib().orc_lnc_table(int(size), _dptr(T))
    return T


def _matrix_ptrs(mats):
    """mats: list (per node) of 2-D arrays or None -> (ctypes array of double*, keepalive)."""
    arr = (_dp * len(mats))()
    keep = []
    for i, m in enumerate(mats):
        if m is None:
            arr[i] = None
        else:
            m = np.ascontiguousarray(m, dtype=np.float64)
            keep.append(m)
            arr[i] = _dptr(m)
    return arr, keep


def node_matrices(tree: FlatTree, lam_per_node, mu_per_node, maxfs):
    """One matrix per non-root node keyed on (int t, lambda, mu) — cafe/cafe_tree.c:374-391,461-483."""
    cache = {}
    mats = []
    for i in range(tree.n_nodes):
        if i == tree.root:
            mats.append(None)
            continue
        key = (int(tree.branchlength[i]), float(lam_per_node[i]), float(mu_per_node[i]))
        if key not in cache:
            cache[key] = bd_matrix(key[0], key[1], key[2], maxfs)
        mats.append(cache[key])
    return mats


def prune(tree: FlatTree, mats, counts_by_leaf, rng, leaf_err=None):
    """rng = (min, max, root_min, root_max).  counts_by_leaf in leaf order (even nlist indices)."""
    S = next(m for m in mats if m is not None).shape[0]
    lc = np.full(tree.n_nodes, -1, dtype=np.int32)
    lc[0::2] = counts_by_leaf
    mp, keep = _matrix_ptrs(mats)
    E = 0
    ep = None
    if leaf_err is not None:
        ep, keep2 = _matrix_ptrs(leaf_err)
        E = next(m for m in leaf_err if m is not None).shape[0]
    out = np.zeros(rng[3] - rng[2] + 1)
    rc = lib().orc_prune(tree.n_nodes, _iptr(tree.left), _iptr(tree.right), tree.root, mp, S, _iptr(lc),
                         ep, E, rng[0], rng[1], rng[2], rng[3], _dptr(out))
    if rc != 0:
        raise ValueError("leaf count outside the likelihood vector")
    return out


def viterbi(tree: FlatTree, mats, counts_by_leaf, rng, leaf_err=None):
    """cafe_tree_viterbi for one family: (sizes per node in nlist order, max root likelihood)."""
    S = next(m for m in mats if m is not None).shape[0]
    lc = np.full(tree.n_nodes, -1, dtype=np.int32)
    lc[0::2] = counts_by_leaf
    mp, keep = _matrix_ptrs(mats)
    E = 0
    ep = None
    if leaf_err is not None:
        ep, keep2 = _matrix_ptrs(leaf_err)
        E = next(m for m in leaf_err if m is not None).shape[0]
    sizes = np.zeros(tree.n_nodes, dtype=np.int32)
    ml = C.c_double(0)
    rc = lib().orc_viterbi(tree.n_nodes, _iptr(tree.left), _iptr(tree.right), tree.root, mp, S, _iptr(lc),
                           ep, E, rng[0], rng[1], rng[2], rng[3], _iptr(sizes), C.byref(ml))
    if rc != 0:
        raise ValueError("leaf count outside the likelihood vector")
    return sizes, ml.value


def forced_range(counts_by_leaf):
    """cafe_family_set_size_with_family_forced (cafe/cafe_family.c:236-255): (min, max, root_min, root_max) of one family."""
    mx = int(max(counts_by_leaf))
    return (0, mx + max(50, mx // 5), 1, int(np.rint(mx * 1.25)))


def viterbi_branch_pvalues(tree: FlatTree, mats, sizes, range_max):
    S = next(m for m in mats if m is not None).shape[0]
    mp, keep = _matrix_ptrs(mats)
    sizes = np.ascontiguousarray(sizes, dtype=np.int32)
    out = np.zeros(tree.n_nodes)
    lib().orc_viterbi_branch_pvalues(tree.n_nodes, _iptr(tree.left), _iptr(tree.right), tree.root, mp, S, _iptr(sizes), range_max, _dptr(out))
    return out


def chi2cdf(x, df=1):
    return lib().orc_chi2cdf(float(x), int(df))


def lrt_family(tree: FlatTree, mats, lam_per_node, mu_per_node, branchlength, counts_by_leaf, rng, leaf_err=None):
    """Branch-stretch likelihood-ratio test of one family (cafe/cafe_main.c:342-396).  `branchlength` (float64 array) is updated in
    place the way the reference's tree copy is (restored through an int).  mu_per_node keys the LENGTHENED branches only (`mats`
    are the tree's own matrices): zeros restate the stock binary (its tree copy carries the tree-level mu = 0), the nodes' own mu
    the algorithm as written.  Returns (ratios, best likelihood, steps) per node."""
    S = next(m for m in mats if m is not None).shape[0]
    lc = np.full(tree.n_nodes, -1, dtype=np.int32)
    lc[0::2] = counts_by_leaf
    mp, keep = _matrix_ptrs(mats)
    E = 0
    ep = None
    if leaf_err is not None:
        ep, keep2 = _matrix_ptrs(leaf_err)
        E = next(m for m in leaf_err if m is not None).shape[0]
    lam = np.ascontiguousarray(lam_per_node, dtype=np.float64)
    mu = np.ascontiguousarray(mu_per_node, dtype=np.float64)
    assert branchlength.dtype == np.float64 and branchlength.flags.c_contiguous
    ratios = np.zeros(tree.n_nodes)
    best = np.zeros(tree.n_nodes)
    steps = np.zeros(tree.n_nodes, dtype=np.int32)
    rc = lib().orc_lrt_family(tree.n_nodes, _iptr(tree.left), _iptr(tree.right), tree.root, mp, S, _dptr(lam), _dptr(mu),
                              _dptr(branchlength), _iptr(lc), ep, E, rng[0], rng[1], rng[2], rng[3], _dptr(ratios), _dptr(best),
                              _iptr(steps))
    if rc != 0:
        raise ValueError("leaf count outside the likelihood vector")
    return ratios, best, steps


def score(tree: FlatTree, mats, counts, rng, prior, ref_idx=None, leaf_err=None, want_L=False):
    counts = np.ascontiguousarray(counts, dtype=np.int32)
    F = counts.shape[0]
    S = next(m for m in mats if m is not None).shape[0]
    mp, keep = _matrix_ptrs(mats)
    E = 0
    ep = None
    if leaf_err is not None:
        ep, keep2 = _matrix_ptrs(leaf_err)
        E = next(m for m in leaf_err if m is not None).shape[0]
    prior = np.ascontiguousarray(prior, dtype=np.float64)
    logpost = np.zeros(F)
    maxlik = np.zeros(F)
    argmax = np.zeros(F, dtype=np.int32)
    rf = rng[3] - rng[2] + 1
    L_all = np.zeros((F, rf)) if want_L else None
    fz = C.c_int(-1)
    refp = None
    if ref_idx is not None:
        ref_idx = np.ascontiguousarray(ref_idx, dtype=np.int32)
        refp = _iptr(ref_idx)
    sc = lib().orc_score(tree.n_nodes, _iptr(tree.left), _iptr(tree.right), tree.root, mp, S, ep, E,
                         rng[0], rng[1], rng[2], rng[3], _dptr(prior), F, _iptr(counts), refp,
                         _dptr(logpost), _dptr(maxlik), _iptr(argmax),
                         _dptr(L_all) if want_L else None, C.byref(fz))
    return {"score": sc, "logpost": logpost, "maxlik": maxlik, "argmax": argmax, "L": L_all,
            "first_zero": fz.value}


def prior_poisson(shift, lam, n=1000):
    return np.array([lib().orc_poisspdf(shift - 1 + i, lam) for i in range(n)])


def random_probabilities(tree, mats, rng_min, rng_max, root_size, trials, uniforms=None, leaf_err=None):
    S = next(m for m in mats if m is not None).shape[0]
    mp, keep = _matrix_ptrs(mats)
    used = C.c_long(0)
    srt = np.zeros(trials)
    uns = np.zeros(trials)
    sizes = np.zeros((trials, tree.n_nodes), dtype=np.int32)
    caps = np.zeros(trials, dtype=np.int32)
    up = None
    if uniforms is not None:
        uniforms = np.ascontiguousarray(uniforms, dtype=np.float64)
        up = _dptr(uniforms)
    E = 0
    ep = None
    if leaf_err is not None:
        ep, keep2 = _matrix_ptrs(leaf_err)
        E = next(m for m in leaf_err if m is not None).shape[0]
    lib().orc_random_probabilities(tree.n_nodes, _iptr(tree.left), _iptr(tree.right), tree.root, mp, S,
                                   rng_min, rng_max, root_size, trials, up, C.byref(used),
                                   _dptr(srt), _dptr(uns), _iptr(sizes), _iptr(caps), ep, E)
    return {"sorted": srt, "unsorted": uns, "sizes": sizes, "caps": caps, "used": used.value}


def pvalue(v, cd):
    cd = np.ascontiguousarray(cd, dtype=np.float64)
    return lib().orc_pvalue(float(v), _dptr(cd), len(cd))


def family_pvalue(tree, mats, counts_by_leaf, cd, leaf_err=None):
    S = next(m for m in mats if m is not None).shape[0]
    mp, keep = _matrix_ptrs(mats)
    lc = np.full(tree.n_nodes, -1, dtype=np.int32)
    lc[0::2] = counts_by_leaf
    cd = np.ascontiguousarray(cd, dtype=np.float64)
    E = 0
    ep = None
    if leaf_err is not None:
        ep, keep2 = _matrix_ptrs(leaf_err)
        E = next(m for m in leaf_err if m is not None).shape[0]
    pv = np.zeros(max(1, cd.shape[0] + 8))
    rf = C.c_int(0)
    p = lib().orc_family_pvalue(tree.n_nodes, _iptr(tree.left), _iptr(tree.right), tree.root, mp, S, ep, E,
                                _iptr(lc), _dptr(cd), cd.shape[0], cd.shape[1], _dptr(pv), C.byref(rf))
    return p, pv[:max(0, min(rf.value, cd.shape[0]))]


# --------------------------------------------------------------------------- test-side data generation

def random_tree(n_leaves: int, seed: int = 1, max_gap: int = 3) -> str:
    """Random ultrametric binary tree with integer branch lengths >= 1 (SURVEY.md 8d recipe)."""
    rng = np.random.RandomState(seed)
    nodes = [(f"s{i}", 0) for i in range(n_leaves)]
    h = 0
    while len(nodes) > 1:
        h += int(rng.randint(1, max_gap + 1))
        i, j = sorted(rng.choice(len(nodes), 2, replace=False))
        a, b = nodes[i], nodes[j]
        new = (f"({a[0]}:{h - a[1]},{b[0]}:{h - b[1]})", h)
        nodes = [x for k, x in enumerate(nodes) if k not in (i, j)] + [new]
    return nodes[0][0]


def simulate_families(tree: FlatTree, lam_per_node, mu_per_node, maxfs, n_families, root_sizes, seed=0):
    """Draw family tables from the model with the ORACLE's matrices (CPU only)."""
    rng = np.random.RandomState(seed)
    mats = node_matrices(tree, lam_per_node, mu_per_node, maxfs)
    cdfs = {id(m): np.cumsum(m, axis=1) for m in mats if m is not None}
    order, st = [], [tree.root]
    while st:
        v = st.pop()
        order.append(v)
        if tree.left[v] >= 0:
            st.append(tree.right[v])
            st.append(tree.left[v])
    sizes = np.zeros((n_families, tree.n_nodes), dtype=np.int64)
    sizes[:, tree.root] = rng.choice(root_sizes, size=n_families)
    for v in order:
        if v == tree.root:
            continue
        cdf = cdfs[id(mats[v])]
        u = rng.random_sample(n_families)
        par = sizes[:, tree.parent[v]]
        child = np.empty(n_families, dtype=np.int64)
        for p in np.unique(par):
            idx = np.where(par == p)[0]
            child[idx] = np.searchsorted(cdf[p], u[idx], side="left")
        sizes[:, v] = np.minimum(child, maxfs)
    return sizes[:, 0::2].astype(np.int32)


def conditional_distribution(tree: FlatTree, mats, ranges, n_samples, uniforms=None, leaf_err=None):
    """cafe/conditional_distribution.cpp:48-57 single-threaded: rows for s = root_min..root_max.
    uniforms (optional): the unifrnd() stream in the reference's order; None draws from glibc rand()."""
    rmin, rmax, root_min, root_max = ranges
    rows = []
    off = 0
    per_row = n_samples * (tree.n_nodes - 1)
    for s in range(root_min, root_max + 1):
        u = None if uniforms is None else uniforms[off:off + per_row]
        r = random_probabilities(tree, mats, rmin, rmax, s, n_samples, u, leaf_err=leaf_err)
        rows.append(r["sorted"])
        off += per_row
    return np.array(rows)


# --------------------------------------------------------------------------- branch cutting

def split_tree(tree: FlatTree, b: int):
    """phylogeny_split_tree (libtree/phylogeny.c:571-614) as cafe_tree_split uses it (cafe/cafe_tree.c:517-527): cut the branch
    above node b.  Returns (rest, sub, rest_orig, sub_orig): the remaining tree and the cut-off subtree as FlatTrees in their own
    infix numbering (tree_build_node_list), and for each the original node id of every new node.  In the remaining tree b's
    parent disappears: its other child takes its place and inherits its branch length on top of its own (:594), or becomes the
    root when the parent was the root (:586-591).  Both roots get branch length -1 (:611-612)."""
    assert b != tree.root
    left, right, parent = tree.left.tolist(), tree.right.tolist(), tree.parent.tolist()
    bl = tree.branchlength.tolist()
    par = parent[b]
    sib = right[par] if left[par] == b else left[par]
    if par == tree.root:
        rest_root = sib
    else:
        bl[sib] += bl[par]
        grand = parent[par]
        if left[grand] == par:
            left[grand] = sib
        else:
            right[grand] = sib
        rest_root = tree.root

    def flatten(root):
        order = []

        def infix(v):
            if left[v] >= 0:
                infix(left[v]); order.append(v); infix(right[v])
            else:
                order.append(v)

        infix(root)
        new = {v: i for i, v in enumerate(order)}
        L = [new[left[v]] if left[v] >= 0 else -1 for v in order]
        R = [new[right[v]] if right[v] >= 0 else -1 for v in order]
        P = [-1] * len(order)
        for i, v in enumerate(order):
            if L[i] >= 0:
                P[L[i]] = i; P[R[i]] = i
        B = [bl[v] for v in order]
        B[new[root]] = -1.0
        return FlatTree(L, R, P, B, [tree.names[v] for v in order]), order

    rest, rest_orig = flatten(rest_root)
    sub, sub_orig = flatten(b)
    return rest, sub, rest_orig, sub_orig


def branch_cut(tree: FlatTree, lam_per_node, mu_per_node, counts, ranges, n_samples, b, max_pvalues, cutoff, uniforms=None):
    """cut_branch + compute_cutpvalues (cafe/branch_cutting.cpp:185-219, :101-150) for branch b and the families `counts`
    (F x n_leaves, leaf order of `tree`).  The conditional distributions draw from glibc rand() (after the caller's srand) or
    from `uniforms`; the remaining tree's distribution first, then the subtree's.  Returns {"pvalues": F values (-1 where
    max_pvalues > cutoff), "cd1", "cd2" (None when one side is a single leaf), "rest", "sub"}."""
    counts = np.asarray(counts)
    lam_per_node = np.asarray(lam_per_node, dtype=np.float64); mu_per_node = np.asarray(mu_per_node, dtype=np.float64)
    if b == tree.root:
        return {"pvalues": np.zeros(len(counts)), "cd1": None, "cd2": None}
    rest, sub, ro, so = split_tree(tree, b)
    maxfs = max(ranges[1], ranges[3])
    m_rest = node_matrices(rest, lam_per_node[ro], mu_per_node[ro], maxfs) if rest.n_nodes > 1 else None
    m_sub = node_matrices(sub, lam_per_node[so], mu_per_node[so], maxfs) if sub.n_nodes > 1 else None
    leaf_col = {v: v // 2 for v in range(0, tree.n_nodes, 2)}            # original leaf id -> column of counts
    cols_rest = [leaf_col[ro[i]] for i in range(0, rest.n_nodes, 2)]
    cols_sub = [leaf_col[so[i]] for i in range(0, sub.n_nodes, 2)]
    off = [0]

    def cd(t, mats, n):
        per = (ranges[3] - ranges[2] + 1) * n * (t.n_nodes - 1)
        u = None if uniforms is None else uniforms[off[0]:off[0] + per]
        off[0] += per
        return conditional_distribution(t, mats, ranges, n, u)

    rf = ranges[3] - ranges[2] + 1
    out = np.zeros(len(counts))
    if sub.n_nodes == 1 or rest.n_nodes == 1:
        t, mats, cols = (rest, m_rest, cols_rest) if sub.n_nodes == 1 else (sub, m_sub, cols_sub)
        cd1, cd2 = cd(t, mats, n_samples), None
        for f in range(len(counts)):
            if max_pvalues[f] > cutoff:
                out[f] = -1.0
                continue
            L = prune(t, mats, counts[f, cols], ranges)
            out[f] = max(pvalue(L[s], cd1[s][:n_samples]) for s in range(rf))
    else:
        n10 = n_samples // 10
        cd1 = cd(rest, m_rest, n10)
        cd2 = cd(sub, m_sub, n10)
        for f in range(len(counts)):
            if max_pvalues[f] > cutoff:
                out[f] = -1.0
                continue
            l1 = prune(rest, m_rest, counts[f, cols_rest], ranges)
            l2 = prune(sub, m_sub, counts[f, cols_sub], ranges)
            best = 0.0
            with np.errstate(divide="ignore", invalid="ignore"):
                for s2 in range(rf):          # p_values_of_two_trees, branch_cutting.cpp:20-44
                    for s1 in range(rf):
                        p = 0.0
                        for t in range(n10):
                            p += pvalue(float(np.float64(l1[s1]) * np.float64(l2[s2]) / np.float64(cd2[s2][t])), cd1[s1][:n10])
                        p = p / n10
                        if p > best:
                            best = p
            out[f] = best
    return {"pvalues": out, "cd1": cd1, "cd2": cd2, "rest": rest, "sub": sub, "rest_orig": ro, "sub_orig": so}


def srand(seed: int):
    C.CDLL(None).srand(C.c_uint(seed))
