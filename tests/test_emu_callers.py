"""SURVEY.md section 8f rows 1-2 on the CPU: the documented callers of the pool (README VisionLanguageModel, the x-ray
fusion model) trained for several steps on the product path under the host emulation (tests/cuda_emu), against their
oracle twins -- the bodies of tests/test_gpu_callers.py on CPU tensors."""
import pytest

from tests import test_gpu_callers as CL
from tests.emu_support import cuda_emulation  # noqa: F401  (fixture)

pytestmark = pytest.mark.usefixtures("cuda_emulation")


@pytest.mark.parametrize("test", [CL.test_vision_language_model_training_steps,
                                  CL.test_xray_fusion_model_missing_modalities_and_curriculum_toggle], ids=lambda f: f.__name__[5:])
def test_caller(monkeypatch, test):
    monkeypatch.setattr(CL, "DEV", "cpu")
    test()
