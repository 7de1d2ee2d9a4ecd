"""The tcgen05 / TMEM / TMA projection GEMM on the CPU: gemm_tcgen05.cu compiled for the host emulation, with the
functional model of tests/cuda_emu/tcgen05_emu.h behind its inline-PTX wrappers (mbarriers, tiled TMA with the 128-byte
swizzle, cluster multicast, tensor memory, tcgen05.mma through the shared-memory / instruction descriptors,
cta_group::2).  The model's layouts are validated by the kernels measured on hardware computing correct products under
it; with that, the variants that have NOT run on hardware yet (AECF_GEMM_EPI=2 / 3, AECF_GEMM_2SM_EW=8, AECF_GEMM_2SM_AUX=1) are
checked for what a functional model can see -- barrier counts and phases (a wrong count deadlocks the emulation), tile /
box / column indexing, staging layout -- not for missing waits or fences, and not for speed.
"""
import os
import subprocess
import sys

import pytest
import torch

from aecf_b200 import _lib
from tests import test_gpu_gemm_tcgen05 as G
from tests import test_gpu_parity as P
from tests.emu_support import cuda_emulation  # noqa: F401  (fixture)
from tests.golden.cases import Case

pytestmark = pytest.mark.usefixtures("cuda_emulation")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# one tile; BN = 256; a cluster of two with multicast; ragged everywhere (zero fill, clipped stores); three 256-wide column
# tiles and three tiles per CTA (both accumulators re-used); cta_group::2 (10 k-blocks); split-K with a ragged k tail
# several row-block groups per CTA pair with two column tiles each (the resident-A-panel kernel's slot hand-over)
SHAPES = [(128, 128, 64), (256, 256, 64), (384, 256, 128), (200, 136, 72), (392, 520, 200), (512, 256, 640), (300, 512, 4100),
          (1024, 512, 128), (512, 512, 520)]         # ... and nine k-blocks, the last one ragged: the dX contraction
LAYOUTS = [(_lib.K_MAJOR, _lib.K_MAJOR), (_lib.K_MAJOR, _lib.MN_MAJOR), (_lib.MN_MAJOR, _lib.MN_MAJOR), (_lib.MN_MAJOR, _lib.K_MAJOR)]


@pytest.fixture(autouse=True)
def _cpu_stands_in_for_the_device(monkeypatch):
    monkeypatch.setattr(G, "DEV", "cpu")
    monkeypatch.setattr(P, "DEV", "cpu")
    plain_to = torch.Tensor.to
    monkeypatch.setattr(torch.Tensor, "to", lambda self, *a, **k: (lambda r: r.clone() if r is self else r)(plain_to(self, *a, **k)))


@pytest.mark.parametrize("layouts", LAYOUTS, ids=["aK_bK", "aK_bMN", "aMN_bMN", "aMN_bK"])
@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_kernels(shape, layouts):
    G.test_tcgen05_gemm(shape, *layouts)


def test_cta_pair_kernels_with_an_odd_number_of_row_blocks():
    """TMA stores of the emulation read their staging box at the latest moment the kernel allows.  Five row blocks on CTA
    pairs leave a filler tile; every epilogue round commits one (possibly empty) bulk group, so `wait_group.read 1` keeps
    protecting the buffer of the previous tile's last store.  (The round-1 instantiation without that commit failed here.)"""
    if os.environ.get("AECF_GEMM_CLUSTER") == "1":
        pytest.skip("without clusters the product runs on the single-CTA kernel")
    G.test_tcgen05_gemm((640, 256, 640), _lib.K_MAJOR, _lib.K_MAJOR)


def test_strided_operands_and_kernel_choice():
    G.test_tcgen05_strided_output_and_operand_views()
    G.test_tcgen05_is_what_auto_picks_for_the_projection_shapes()


@pytest.mark.parametrize("shape", [(640, 512, 512, 8), (300, 256, 128, 4)], ids=lambda s: "x".join(map(str, s)))
def test_side_output(shape):
    P.test_gemm_with_side_output(shape, torch.bfloat16)


def test_whole_step_on_the_tensor_core_gemms():
    """B*M = 480 rows: the folded forward (192-wide tiles, fp32 score side output), the K = D + 8 contraction of dX and the
    split-K weight gradients all take the tcgen05 kernels, against the stage-rounded oracle with bit-exact masks."""
    case = Case("emu_d256_h8_m3_b160", B=160, M=3, D=256, H=8, dropout=0.1, pooled_grad=True, data_seed=77, offset=4)
    P.test_bf16_masks_exact_against_stage_rounded_oracle(case, True)
    P.test_bf16_masks_exact_against_stage_rounded_oracle(case, False)


def test_reported_kernel_follows_the_switches():
    """aecf_gemm_last_kernel() names what ran, so an A/B run (and the variant runs below) can check that a switch took."""
    from aecf_b200 import ops
    env = os.environ
    ew = "8" if env.get("AECF_GEMM_2SM_EW") == "8" else "4"
    epi = env.get("AECF_GEMM_EPI", "1")
    cluster = "1" if env.get("AECF_GEMM_CLUSTER") == "1" else "2"
    bf = torch.bfloat16
    a, b = torch.randn(512, 128, dtype=bf), torch.randn(256, 128, dtype=bf)
    apanel = env.get("AECF_GEMM_APANEL") == "1" and cluster == "2"
    ops.gemm(a, b, m=512, n=256, k=128, a_layout=_lib.K_MAJOR, b_layout=_lib.K_MAJOR, lda=128, ldb=128)           # 2 k-blocks
    assert _lib.gemm_last_kernel() == (f"tcgen05 apanel bn256 ew{ew} kb8" if apanel else f"tcgen05 1sm bn256 cluster{cluster} epi{epi} splits1")
    ops.gemm(a, b, m=512, n=256, k=128, a_layout=_lib.K_MAJOR, b_layout=_lib.K_MAJOR, lda=128, ldb=128, out_dtype=torch.float32)
    assert _lib.gemm_last_kernel() == (f"tcgen05 apanel bn256 ew{ew} kb8" if apanel else
                                       f"tcgen05 1sm bn256 cluster{cluster} epi{'1' if epi == '3' else epi} splits1")   # EPI 3: bf16 output only
    a, b = torch.randn(512, 640, dtype=bf), torch.randn(256, 640, dtype=bf)
    ops.gemm(a, b, m=512, n=256, k=640, a_layout=_lib.K_MAJOR, b_layout=_lib.K_MAJOR, lda=640, ldb=640)           # 10 k-blocks
    assert _lib.gemm_last_kernel() == (f"tcgen05 2sm bn256 ew{ew} splits1" if cluster == "2" else f"tcgen05 1sm bn256 cluster1 epi{epi} splits1")
    a, b = torch.randn(640, 512, dtype=bf), torch.randn(520, 512, dtype=bf)
    ops.gemm_aux(a, b, m=640, n=512, k=512, aux_cols=8)
    want = f"tcgen05 2sm bn192 ew{ew} aux splits1" if (env.get("AECF_GEMM_2SM_AUX") == "1" and cluster == "2") else f"tcgen05 1sm bn192 cluster{cluster} epi{epi} splits1"
    if apanel:
        want = f"tcgen05 apanel bn192 ew{ew} kb8"                              # K = 512: eight k-blocks, the panel fits
    assert _lib.gemm_last_kernel() == want
    ops.gemm(a[:64].float(), b[:64].float(), m=64, n=64, k=512, a_layout=_lib.K_MAJOR, b_layout=_lib.K_MAJOR, lda=512, ldb=512)
    assert _lib.gemm_last_kernel() == "simt"


VARIANTS = {"pipelined_epilogue_and_fixed_cta_pairs": {"AECF_GEMM_EPI": "2", "AECF_GEMM_2SM_FIX": "1"},
            "eight_warp_epilogues": {"AECF_GEMM_EPI": "3", "AECF_GEMM_2SM_EW": "8"},          # 1SM (cluster of two) and cta_group::2
            "eight_warp_epilogue_no_cluster": {"AECF_GEMM_EPI": "3", "AECF_GEMM_CLUSTER": "1"},
            "score_columns_on_cta_pairs": {"AECF_GEMM_2SM_AUX": "1", "AECF_GEMM_2SM_EW": "8"},
            "resident_a_panel": {"AECF_GEMM_APANEL": "1", "AECF_GEMM_2SM_EW": "8"}}     # four epilogue warps: checked by hand runs


@pytest.fixture(scope="module")
def variant_runs():
    """The library reads these switches once per process, hence one child pytest per variant, running the tests above with
    the switches set; all children are started together and each test below waits for its own."""
    if os.environ.get("AECF_EMU_GEMM_CHILD") == "1":
        yield {}
        return
    from tests.emu_support import load_emulation
    load_emulation()                                     # build once, before the children race for it
    select = "(test_kernels and (384x256x128 or 392x520x200 or 512x256x640 or 1024x512x128 or 512x512x520)) or side_output or whole_step or reported_kernel or odd_number_of_row_blocks"
    runs = {}
    for name, switches in VARIANTS.items():
        env = dict(os.environ, AECF_EMU_GEMM_CHILD="1", **switches)
        runs[name] = subprocess.Popen([sys.executable, "-m", "pytest", "-q", "-x", "-p", "no:cacheprovider", os.path.abspath(__file__),
                                       "-k", select], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env, cwd=ROOT)
    yield runs
    for proc in runs.values():
        if proc.poll() is None:
            proc.kill()


@pytest.mark.parametrize("variant", sorted(VARIANTS))
def test_epilogue_variant(variant_runs, variant):
    if os.environ.get("AECF_EMU_GEMM_CHILD") == "1":
        pytest.skip("already inside a variant run")
    out, _ = variant_runs[variant].communicate(timeout=1500)
    assert variant_runs[variant].returncode == 0, out[-4000:]
