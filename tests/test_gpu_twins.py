"""GPU twins of the cross-model cases tests/test_host_logic.py runs on the CPU test double: the same bodies (fit,
transform / predict / inverse_transform, the rotators of plain and whitened models) against the same oracle
restatements, with the model classes on the real CudaOps (the C-ABI kernels)."""
import pytest

import test_host_logic as H

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _real_ops(monkeypatch):
    monkeypatch.setattr(H, "make_ops", lambda: None)  # ops=None: the classes build their own CudaOps


@pytest.mark.parametrize("cls,alpha", [("CCA", (0.0, 0.0)), ("RDA", (0.0, 1.0)), ("CPCCA", 0.2)])
def test_cpcca_family_and_whitened_rotator(cls, alpha):
    H.test_cpcca_family_host_logic(cls, alpha)


@pytest.mark.parametrize("mode", ["implicit", "pca", "cca"])
def test_cross_transform_predict_inverse(mode):
    H.test_cross_transform_predict_inverse_host_logic(mode)


def test_mca_rotator_and_its_transform():
    H.test_mca_rotator_host_logic()


def test_mca_more_samples_than_features_default_pca():
    H.test_mca_more_samples_than_features_default_pca()


@pytest.mark.parametrize("use_pca", [True, False])
def test_mca_variance_based_n_modes(use_pca):
    H.test_mca_variance_based_n_modes(use_pca)


def test_inverse_transform_rejects_modes_the_model_does_not_hold():
    H.test_inverse_transform_rejects_modes_the_model_does_not_hold()


def test_rotator_compute_false_runs_max_iter_without_raising():
    H.test_rotator_compute_false_runs_max_iter_without_raising()
