"""GPU tests of the tcgen05/TMEM/TMA MLP kernel, each in its own process."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def run_check(variant, rows, seed=0, out_div=1.0):
    r = subprocess.run([sys.executable, "-m", "tests.tc_check", "--variant", str(variant), "--rows", str(rows),
                        "--seed", str(seed), "--out-div", str(out_div)],
                       cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, f"tc_check crashed:\n{r.stdout[-2000:]}\n{r.stderr[-3000:]}"
    return json.loads(r.stdout.strip().splitlines()[-1])


@pytest.mark.parametrize("rows", [1, 128, 1000, 40000])
def test_cta_pair_kernel(rows):
    """Default kernel: CTA pair, tcgen05.mma.cta_group::2."""
    res = run_check(2, rows, seed=rows)
    assert res["finite"]
    assert res["max_err"] <= 4e-3 * max(1.0, res["ref_absmax"]), res


@pytest.mark.parametrize("rows", [129, 5000])
def test_single_cta_kernel(rows):
    """Bring-up variant (LIST_B200_MLP_VARIANT=1): cta_group::1."""
    res = run_check(1, rows, seed=rows)
    assert res["finite"]
    assert res["max_err"] <= 4e-3 * max(1.0, res["ref_absmax"]), res


def test_out_div_is_a_true_division():
    res = run_check(2, 300, seed=7, out_div=10.0)
    assert res["max_err"] <= 4e-4 * max(1.0, res["ref_absmax"] * 10), res


@pytest.mark.parametrize("variant", [1, 2])
def test_hidden_activations_layer_by_layer(variant):
    r = subprocess.run([sys.executable, "-m", "tests.tc_check", "--variant", str(variant), "--rows", "700", "--layers"],
                       cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, f"tc_check crashed:\n{r.stdout[-2000:]}\n{r.stderr[-3000:]}"
    res = json.loads(r.stdout.strip().splitlines()[-1])
    print(res)
    assert res["h1_err"] <= 1e-3 * max(1.0, res["h1_absmax"]), res
    assert res["h2_err"] <= 2e-2 and res["h3_err"] <= 2e-2, res
    assert res["debug_equals_plain"]
