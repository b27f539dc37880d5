"""CUDA path vs the reference-generated golden vectors (bit-exact), through the C ABI."""
import pytest

import golden_io

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("group", golden_io.GROUPS)
def test_cuda_reproduces_reference_golden_vectors(gpu, group):
    assert golden_io.check(lambda p: gpu.solve(p)[0], group) > 0


@pytest.mark.parametrize("group", golden_io.GROUPS)
def test_column_kernel_reproduces_golden_vectors(gpu, group, monkeypatch):
    monkeypatch.setenv("WHMEC_FORCE_COLUMN_KERNEL", "1")
    assert golden_io.check(lambda p: gpu.solve(p)[0], group) > 0
