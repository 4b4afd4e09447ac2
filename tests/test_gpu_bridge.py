"""The reference with integration/gpu_bridge.cpp compiled in (oracle/Makefile target `bridge`: every reference source
unmodified except the two statements INTEGRATION.md §2 names) — proof that the C-ABI drops into CAFE itself.

not gpu: the binary exists, is linked against libcafe_gpu.so, and without a CUDA device every objective call fails loudly
         (no CPU fallback: the search sees -inf everywhere).
-m gpu : its `lambda -s` transcript on example/example_data.tab follows the stock binary's call for call
         (tests/golden/example.npz, written from the unmodified binary by tests/golden/make_golden.py)."""
import os
import subprocess

import numpy as np
import pytest

import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
EX_TREE = "(((chimp:6,human:6):81,(mouse:17,rat:17):70):6,dog:93)"


def run_bridge(tmp_path, command):
    z = np.load(os.path.join(GOLD, "example.npz"))
    species = [str(s) for s in z["species_leaf_order"]]
    with open(tmp_path / "example_data.tab", "w") as f:
        f.write("\t".join(["FAMILYDESC", "FAMILY"] + species) + "\n")
        for i, r in zip(z["ids"], z["counts"]):
            f.write("\t".join(["d", str(i)] + [str(x) for x in r]) + "\n")
    (tmp_path / "s.sh").write_text("seed 10\nload -i example_data.tab -t 1\ntree %s\n%s\n" % (EX_TREE, command))
    r = subprocess.run([oracle.ref_gpu_binary(), "s.sh"], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    return z, r


needs_bridge = pytest.mark.skipif("oracle.ref_gpu_binary() is None", reason="oracle/_ref/cafe_ref_gpu not built (needs /root/reference)")


@needs_bridge
def test_bridge_binary_links_the_c_abi_library():
    out = subprocess.run(["ldd", oracle.ref_gpu_binary()], capture_output=True, text=True).stdout
    line = [ln for ln in out.splitlines() if "libcafe_gpu.so" in ln]
    assert line and "not found" not in line[0]


@needs_bridge
def test_bridge_has_no_cpu_fallback(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    z, r = run_bridge(tmp_path, "lambda -l 0.005 -score")
    assert "no CPU fallback" in r.stderr
    assert "Score: -inf" in r.stdout


@needs_bridge
@pytest.mark.gpu
def test_bridge_lambda_search_transcript_equals_the_stock_binary(tmp_path):
    z, r = run_bridge(tmp_path, "lambda -s")
    assert r.returncode == 0, r.stderr[-2000:]
    ours = []
    for ln in r.stdout.splitlines():
        ln = ln.lstrip(".")
        if ln.startswith("Lambda : ") and "& Score:" in ln:
            a, b = ln[len("Lambda : "):].split(" & Score: ")
            ours.append((float(a), float(b)))
    ref = z["search_trace"]
    assert len(ours) == len(ref) + 1                                 # + the result line
    for (lam, score), (rlam, rscore) in zip(ours, ref):
        assert lam == rlam                                           # same vertex, to the printed digits
        assert abs(score - rscore) <= 1e-6 or (np.isinf(score) and np.isinf(rscore))
    assert abs(ours[-1][0] - float(z["search_lambda"])) <= 1e-6 * float(z["search_lambda"])
    assert abs(ours[-1][1] - float(z["search_neg_score"])) <= 1e-6
