"""Opt-in variants that have NOT yet been run on hardware (round-2 candidates, profiles/r1_gemm_experiments.md).
Skipped unless AECF_TEST_EXPERIMENTAL=1: the driver's `pytest -m gpu` must only see measured code paths.
Each variant is selected by an environment variable the library reads once per process, hence the subprocesses.

    AECF_TEST_EXPERIMENTAL=1 python -m pytest tests/test_gpu_experimental.py -m gpu -q
"""
import os
import subprocess
import sys

import pytest

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("AECF_TEST_EXPERIMENTAL") != "1", reason="opt-in: AECF_TEST_EXPERIMENTAL=1")]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _pytest_with(env, selection):
    e = dict(os.environ)
    e.update(env)
    e.pop("AECF_TEST_EXPERIMENTAL", None)
    res = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", "-p", "no:cacheprovider", "--timeout", "300"]
                         + selection, capture_output=True, text=True, timeout=1500, env=e, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-1000:]


def test_pipelined_gemm_epilogue_passes_the_gemm_and_parity_suites():
    """AECF_GEMM_EPI=2: next tcgen05.ld in flight during the conversion, staging boxes alternating with wait_group.read 1."""
    _pytest_with({"AECF_GEMM_EPI": "2"},
                 ["tests/test_gpu_gemm_tcgen05.py", "tests/test_gpu_parity.py", "-k",
                  "gemm or side_output or bf16 or folded or headline or sharding"])


def test_eight_warp_gemm_epilogue_passes_the_gemm_and_parity_suites():
    """AECF_GEMM_EPI=3: eight epilogue warps (two per TMEM lane quadrant), direct bf16 output only."""
    _pytest_with({"AECF_GEMM_EPI": "3"},
                 ["tests/test_gpu_gemm_tcgen05.py", "tests/test_gpu_parity.py", "-k",
                  "gemm or side_output or bf16 or folded or headline or sharding"])


def test_eight_warp_epilogue_of_the_cta_pair_kernel_passes_the_gemm_and_parity_suites():
    """AECF_GEMM_2SM_EW=8: the cta_group::2 kernel with two epilogue warps per TMEM lane quadrant (gemm_tcgen05_2sm.inc)."""
    _pytest_with({"AECF_GEMM_2SM_EW": "8"},
                 ["tests/test_gpu_gemm_tcgen05.py", "tests/test_gpu_parity.py", "-k",
                  "gemm or side_output or bf16 or folded or headline or sharding"])


def test_streaming_pool_backward_passes_the_parity_and_graph_suites():
    """AECF_POOL_BWD_STREAM=1: the folded backward with cp.async-staged rows (pool_bwd_stream_kernel)."""
    _pytest_with({"AECF_POOL_BWD_STREAM": "1"}, ["tests/test_gpu_parity.py", "tests/test_gpu_graphs.py", "-k", "bf16 or folded or headline or sharding or graph"])


def test_side_output_product_on_cta_pairs_passes_the_gemm_and_parity_suites():
    """AECF_GEMM_2SM_AUX=1: the folded forward's 192-wide tiles with the fp32 score side output on the cta_group::2 kernel."""
    _pytest_with({"AECF_GEMM_2SM_AUX": "1"},
                 ["tests/test_gpu_gemm_tcgen05.py", "tests/test_gpu_parity.py", "-k",
                  "gemm or side_output or bf16 or folded or headline or sharding"])


def test_resident_a_panel_kernel_passes_the_gemm_and_parity_suites():
    """AECF_GEMM_APANEL=1: CTA pairs that keep their 128 x K panel of A in shared memory over all column tiles (K <= 512)."""
    _pytest_with({"AECF_GEMM_APANEL": "1"},
                 ["tests/test_gpu_gemm_tcgen05.py", "tests/test_gpu_parity.py", "-k",
                  "gemm or side_output or bf16 or folded or headline or sharding"])


def test_cta_pair_kernel_with_the_bulk_group_fix_passes_the_gemm_and_parity_suites():
    """AECF_GEMM_2SM_FIX=1: the measured cta_group::2 kernel committing one bulk group per epilogue round, stored or not."""
    _pytest_with({"AECF_GEMM_2SM_FIX": "1"},
                 ["tests/test_gpu_gemm_tcgen05.py", "tests/test_gpu_parity.py", "-k",
                  "gemm or side_output or bf16 or folded or headline or sharding"])
